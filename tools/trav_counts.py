#!/usr/bin/env python
"""Executed node / triangle steps and launch time of one workload under each traversal (pt_scene_set_traversal 2, 3, 4).
    python tools/trav_counts.py --workload c4"""
import argparse, os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from acgpathtracing_b200 import api

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--travs", default="2,3,4")
ap.add_argument("--tess", type=int, default=0, help="use scenes.tessellated_scene(N) instead of the workload's scene")
ap.add_argument("--spp", type=int, default=0)
a = ap.parse_args()
scene_name, w, h, spp, md, nee, imp, desc = bench.WORKLOADS[a.workload]
torch.cuda.set_device(0)
spp = a.spp or spp
if a.tess:
    from acgpathtracing_b200 import scenes
    arr = scenes.tessellated_scene(a.tess)
    print("tessellated_scene", a.tess, "triangles", arr.n_tri, flush=True)
    scene = api.Scene(arr)
else:
    scene = api.Scene(bench.make_scene_arrays(scene_name))
acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
for trav in [int(x) for x in a.travs.split(",")]:
    scene.set_traversal(trav)
    for counted in (0, 2):
        scene.enable_counters(counted)
        if counted:
            scene.counters(reset=True)
        ms = []
        for f in range(3):
            p = api.make_params(w, h, spp, md, frame=f, nee=nee, importance=imp, scene=scene)
            p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            api.launch(p, torch.cuda.current_stream().cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        c = scene.counters(reset=True) if counted else None
        print(f"{a.workload} traversal {trav} {scene.last_launch_info()['traversal']:12s} counted={counted} ms {['%.1f' % m for m in ms]}",
              ({k: v for k, v in c.items() if v} if c else ""), flush=True)
scene.enable_counters(0)
