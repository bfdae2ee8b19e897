"""Generate tests/golden/ref_renders.npz from the reference's OWN programs (oracle/_ref/libptref.so, i.e.
/root/reference/PathTracer_Optix/pathTracerPrograms.cu compiled for the host through oracle/shim/optix.h).

Run in the build container (where /root/reference is mounted):   python tests/golden/make_golden.py
The fixtures travel with the repo, so the oracle port can be pinned on machines without the reference.
Intersection is brute force over all triangles (oracle_geom.h); jitter draw order is the host build's
(recorded in the file as `jitter_order`)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle_api as O  # noqa: E402
from acgpathtracing_b200 import scenes as S  # noqa: E402

# name, scene, w, h, spp, maxDepth, NEE, importance, launches
CASES = [
    ("c1_cornell_is", "cornell", 64, 64, 1, 3, False, True, 1),
    ("c2_cornell_nee", "cornell", 96, 54, 4, 16, True, True, 2),
    ("cornell_uniform", "cornell", 48, 48, 2, 4, False, False, 1),
    ("c3_material_nee", "material", 64, 64, 4, 28, True, True, 2),
    ("material_uniform", "material", 48, 48, 2, 6, False, False, 1),
]


def main():
    port = O.load_port()
    ref = O.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libptref.so is missing: the reference checkout is required")
    scenes = {"cornell": S.cornell32(), "material": S.material_scene(2)}
    out = {"jitter_order": np.int32(ref.ptref_jitter_order())}
    for name, sname, w, h, spp, md, nee, imp, frames in CASES:
        acc = None
        for f in range(frames):
            p = O.make_params(port, w, h, spp, md, frame=f, nee=nee, importance=imp)
            r = O.ref_render(ref, scenes[sname], p, accum=acc, use_bvh=False)
            acc = r["accum"]
        out[name + "/accum"] = r["accum"]
        out[name + "/fb"] = r["fb"]
        out[name + "/cfg"] = np.array([w, h, spp, md, int(nee), int(imp), frames], np.int32)
        out[name + "/rays"] = np.array([r["counters"]["n_radiance"], r["counters"]["n_shadow"]], np.uint64)
        out[name + "/scene"] = np.array(sname)
    np.savez_compressed(os.path.join(HERE, "ref_renders.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_renders.npz"))


if __name__ == "__main__":
    main()
