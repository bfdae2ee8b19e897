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
def cmp(name, acc, o):
    bad = (acc.cpu().numpy().view(np.uint32) != o.view(np.uint32)).any(axis=-1)
    print(name, "differing pixels", int(bad.sum()), np.argwhere(bad)[:4].tolist(), flush=True)
for frames in (1, 3):
    acc_o = None
    for f in range(frames):
        p = O.make_params(port, w, h, spp, md, frame=f, nee=True, importance=True)
        o = osc.render(p, accum=acc_o, math_mode=1, use_bvh=True); acc_o = o["accum"]
    for tune in ("", "team=0"):
        gs = api.Scene(arr); gs.set_tune(tune)
        acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0"); fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
        for f in range(frames):
            p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
            p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
            api.launch(p, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        cmp(f"sequential frames={frames} tune={tune!r}", acc, acc_o)
        gs.destroy()
        group = api.Group.local([0]); gs = group.create_scene(arr); gs.set_tune(tune)
        acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0"); fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
        p = api.make_params(w, h, spp, md, frame=0, nee=True, importance=True, scene=gs)
        p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
        group.launch(p, frames, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        cmp(f"group      frames={frames} tune={tune!r}", acc, acc_o)
        group.destroy()
