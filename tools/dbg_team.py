"""Debug aid: one small launch under several tune settings, compared with the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import oracle_api as O
from acgpathtracing_b200 import api, scenes

port = O.load_port()
arr = scenes.cornell32()
gs, osc = api.Scene(arr), O.OracleScene(port, arr)
w, h, spp, md = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (203, 117, 3, 12)))
p = api.make_params(w, h, spp, md, frame=0, nee=True, importance=True, scene=gs)
o = osc.render(p, math_mode=1, use_bvh=True)
for tune in ("team=0", "", "team=1,team_min=1", "paths32=4,block=128"):
    gs.set_tune(tune)
    for rep in range(3):
        try:
            g = gs.launch_host(p)
            bad = (g["accum"].view(np.uint32) != o["accum"].view(np.uint32)).any(axis=-1)
            print(f"tune={tune!r} rep={rep} launch={gs.last_launch_info()} differing pixels={int(bad.sum())}", "first", np.argwhere(bad)[:3].tolist(), flush=True)
        except Exception as e:
            print(f"tune={tune!r} rep={rep} ERROR {e}", flush=True)
            break
