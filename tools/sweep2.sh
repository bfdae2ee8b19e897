cd /root/repo
python - <<'PY'
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from acgpathtracing_b200 import api, scenes
from acgpathtracing_b200.types import PathTraceParams
arr = scenes.material_scene(3)
gs = api.Scene(arr)
w, h = 1920, 1080
acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0"); fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
tx = (w + 7) // 8
for f in range(3):
    p = api.make_params(w, h, 16, 28, frame=f, nee=True, importance=True, scene=gs)
    q = PathTraceParams.from_buffer_copy(p); q.accumulationBuffer, q.frameBuffer = acc.data_ptr(), fb.data_ptr()
    api.launch(q, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
    o = gs.tile_order()
    print("ORDER frame", f, "n", len(o), "first tiles (x,y):", [(int(t % tx) * 8, int(t // tx) * 4) for t in o[:8]], "unique", len(np.unique(o)) == len(o))
PY
run() { echo "== $1 PTB_TUNE=$2"; PTB_TUNE=$2 timeout 300 python bench.py --workload $1 --steps $3 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run c3 feedback=0 4
run c3 feedback=1 4
run c2 feedback=0 8
run c2 feedback=1 8
run c4 feedback=0 2
run c4 feedback=1 2
