cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ploc or c4 or soups or sah or lbvh" > gpurun_out/pytest_ploc2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ploc2.log
tail -3 gpurun_out/pytest_ploc2.log
python - <<'PY'
import sys, time
sys.path.insert(0, '/root/repo')
from acgpathtracing_b200 import api, scenes
for name, arr in (("cornell", scenes.cornell32()), ("material", scenes.material_scene(3)), ("tess300", scenes.tessellated_scene(300))):
  for rep in range(2):
    t0 = time.time(); gs = api.Scene(arr); dt = time.time() - t0
    info = gs.info(); p = gs.bvh_export_ploc()
    print(f"BUILD {name} rep {rep} tris {info['n_tris']} gpu build_ms {info['build_ms']:.3f} (ploc {p['build_ms']:.3f}, depth {p['depth']}) wall create {dt*1e3:.1f} ms")
    gs.destroy()
PY
