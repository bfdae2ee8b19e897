import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import oracle_api as O
from acgpathtracing_b200 import api, scenes
port = O.load_port()
arr = scenes.cornell32()
osc = O.OracleScene(port, arr)
w, h, spp, md = 203, 117, 3, 12
torch.cuda.set_device(0)
p0 = O.make_params(port, w, h, spp, md, frame=0, nee=True, importance=True)
o = osc.render(p0, math_mode=1, use_bvh=True)["accum"]
def cmp(name, acc):
    bad = (acc.view(np.uint32) != o.view(np.uint32)).any(axis=-1)
    print(name, "differing pixels", int(bad.sum()), np.argwhere(bad)[:4].tolist(), flush=True)
for rep in range(3):
    gs = api.Scene(arr)
    p = api.make_params(w, h, spp, md, frame=0, nee=True, importance=True, scene=gs)
    cmp("fresh scene, launch_host, 1st", gs.launch_host(p)["accum"])
    cmp("same scene, launch_host, 2nd", gs.launch_host(p)["accum"])
    acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0"); fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
    api.launch(p, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    cmp("same scene, device buffers, 3rd", acc.cpu().numpy())
    p.frameBuffer = None
    api.launch(p, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    cmp("same scene, device buffers, no fb", acc.cpu().numpy())
    gs.destroy()
