import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ctypes as C
import numpy as np, torch
from acgpathtracing_b200 import api, scenes
arr = scenes.cornell32()
w, h, spp, md = 203, 117, 3, 12
torch.cuda.set_device(0)
for rep in range(4):
    gs = api.Scene(arr); gs.set_tune("team=2")
    acc = torch.full((h, w, 4), float("nan"), dtype=torch.float32, device="cuda:0")
    p = api.make_params(w, h, spp, md, frame=0, nee=True, importance=True, scene=gs)
    p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), None
    api.launch(p, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    a = (C.c_uint64 * 16)()
    api.lib.pt_scene_get_sched_stats(gs._h, a, 1)
    print("unwritten pixels", int(torch.isnan(acc[..., 0]).sum()), "open teams", a[11], "take %x" % a[12], "commit", a[13], "marks/lock/idle %x" % a[14], flush=True)
    gs.destroy()
