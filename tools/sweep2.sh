cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lbvh or ploc or wide_nodes or sah or c4 or traversal_bit" > gpurun_out/pytest_scan.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_scan.log
tail -5 gpurun_out/pytest_scan.log
python - <<'PY'
import sys, time
sys.path.insert(0, '/root/repo')
from acgpathtracing_b200 import api, scenes
for name, arr in (("cornell", scenes.cornell32()), ("material", scenes.material_scene(3)), ("tess300", scenes.tessellated_scene(300)), ("tess1000", scenes.tessellated_scene(1000))):
    for rep in range(2):
        t0 = time.time(); gs = api.Scene(arr); dt = time.time() - t0
        info = gs.info(); p = gs.bvh_export_ploc()
        print(f"BUILD {name} tris {info['n_tris']} gpu build_ms {info['build_ms']:.3f} (ploc {p['build_ms']:.3f}, depth {p['depth']}) wall create {dt*1e3:.1f} ms tree {gs.tree_in_use()}")
        gs.destroy()
PY
