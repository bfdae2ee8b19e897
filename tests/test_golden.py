"""The oracle port against committed fixtures produced by the reference's own programs
(tests/golden/make_golden.py).  Runs anywhere -- no reference checkout needed."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_api as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_renders.npz")


def _cases():
    z = np.load(GOLD)
    return sorted({k.split("/")[0] for k in z.files if "/" in k})


@pytest.mark.parametrize("case", _cases())
@pytest.mark.parametrize("use_bvh", [False, True])
def test_port_matches_golden(port, cornell, material_scene, case, use_bvh):
    z = np.load(GOLD)
    w, h, spp, md, nee, imp, frames = (int(v) for v in z[case + "/cfg"])
    scene = cornell if str(z[case + "/scene"]) == "cornell" else material_scene
    osc = O.OracleScene(port, scene)
    port.ptor_scene_set_jitter_order(C.c_void_p(osc.h), int(z["jitter_order"]))
    acc = None
    for f in range(frames):
        p = O.make_params(port, w, h, spp, md, frame=f, nee=bool(nee), importance=bool(imp))
        r = osc.render(p, accum=acc, math_mode=0, use_bvh=use_bvh)
        acc = r["accum"]
    assert np.array_equal(r["accum"].view(np.uint32), z[case + "/accum"].view(np.uint32))
    assert np.array_equal(r["fb"], z[case + "/fb"])
    assert r["counters"]["n_radiance"] == int(z[case + "/rays"][0])
    assert r["counters"]["n_shadow"] == int(z[case + "/rays"][1])
