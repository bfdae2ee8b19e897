#!/usr/bin/env python
"""Minimal driver for ncu: N launches of one workload (frames 0..N-1) into device buffers, nothing else.
    python tools/prof_one.py --workload c2 --launches 4 [--tune "team=0"] [--emulate-ranks 8]
Run from the root of the tree whose library is to be profiled (its own bench.py supplies the workload table)."""
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from acgpathtracing_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--tune", default="")
ap.add_argument("--emulate-ranks", type=int, default=0)
ap.add_argument("--spp", type=int, default=0)
a = ap.parse_args()
scene_name, w, h, spp, md, nee, imp, desc = bench.WORKLOADS[a.workload]
spp = a.spp or spp
torch.cuda.set_device(0)
scene = api.Scene(bench.make_scene_arrays(scene_name))
if a.tune and hasattr(scene, "set_tune"):
    scene.set_tune(a.tune)
acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for f in range(a.launches):
    p = api.make_params(w, h, spp, md, frame=f, nee=nee, importance=imp, scene=scene)
    p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
    if a.emulate_ranks > 1:
        api.launch_partition(p, 0, a.emulate_ranks, None)
    else:
        api.launch(p, None)
torch.cuda.synchronize()
print("ok", desc, scene.last_launch_info())
