cd /root/repo
python - <<'PY'
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_int()
for name, a in (("MaxPersistingL2CacheSize", 108), ("MaxAccessPolicyWindowSize", 109), ("L2CacheSize", 38)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), a, 0); print("ATTR", name, v.value)
PY
run() { echo "== $1 PTB_TUNE=$2"; PTB_TUNE=$2 timeout 300 python bench.py --workload $1 --steps $3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('RESULT', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['config']['pipeline'])
    else: print(l.rstrip()[-300:])
"; }
run c4 l2_mb=0 2
run c4 l2_mb=16 2
run c4 l2_mb=32 2
run c4 l2_mb=64 2
run c4 l2_mb=96 2
