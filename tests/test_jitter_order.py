"""Which jitter component takes the first random draw?

`const float2 subpixel_jitter = make_float2(rnd(seed), rnd(seed));` (pathTracerPrograms.cu:730) leaves the order of the two
draws to the compiler.  The reference is device code, so what counts is what NVCC does: this test compiles that very
statement for sm_100a with a constant seed, lets the compiler fold it, and reads the two constants out of the PTX -- the
first draw lands in .x.  The CUDA path (pt_shade.cuh generate_camera_ray) and the oracle's default (jitter_order 0) follow
that; g++ evaluates the arguments right to left, which is why the host build of the reference (oracle/_ref) is pinned with
jitter_order 1 (tests/test_oracle_vs_ref.py).  Both facts are asserted here instead of taken from the survey."""
import os
import shutil
import struct
import subprocess

import pytest

REF_RANDOM_H = "/root/reference/cuda/random.h"
SEED = 1576399551                     # tea<4>(0, 0), SURVEY.md Appendix C
# cuda/random.h:49-67 restated, for machines without the reference checkout
RESTATED = """
static __host__ __device__ __inline__ unsigned int lcg(unsigned int& prev)
{
    const unsigned int LCG_A = 1664525u;
    const unsigned int LCG_C = 1013904223u;
    prev = (LCG_A * prev + LCG_C);
    return prev & 0x00FFFFFF;
}
static __host__ __device__ __inline__ float rnd(unsigned int& prev) { return ((float)lcg(prev) / (float)0x01000000); }
"""
BODY = """
    unsigned int seed = %du;
    const float2 subpixel_jitter = make_float2(rnd(seed), rnd(seed));      // pathTracerPrograms.cu:730, verbatim
    out[0] = subpixel_jitter.x;
    out[1] = subpixel_jitter.y;
""" % SEED


def _draws():
    s, out = SEED, []
    for _ in range(2):
        s = (1664525 * s + 1013904223) & 0xFFFFFFFF
        out.append(struct.unpack("<I", struct.pack("<f", (s & 0xFFFFFF) / float(1 << 24)))[0])
    return out


def _header():
    return f'#include "{REF_RANDOM_H}"\n' if os.path.exists(REF_RANDOM_H) else RESTATED


def test_nvcc_draws_jitter_x_first(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "probe.cu"
    src.write_text('#include <cuda_runtime.h>\n' + _header() + 'extern "C" __global__ void probe(float* out)\n{' + BODY + "}\n")
    ptx = tmp_path / "probe.ptx"
    r = subprocess.run([nvcc, "-arch=sm_100a", "-ptx", "-o", str(ptx), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    consts, stores = {}, {}
    for line in ptx.read_text().splitlines():
        t = line.replace(",", " ").replace(";", " ").split()
        if len(t) >= 3 and t[0].startswith("mov.") and t[2].lstrip("-").isdigit():
            consts[t[1]] = int(t[2]) & 0xFFFFFFFF
        if len(t) >= 3 and t[0].startswith("st.global"):
            off = 4 if "+4]" in t[1] else 0
            stores[off] = consts.get(t[2])
    first, second = _draws()
    assert first == 0x3E96C224 and second == 0x3F320D49            # 0.294449925, 0.695515215 (Appendix C)
    assert stores == {0: first, 4: second}, (stores, "nvcc no longer folds the probe, or the order changed")


def test_gxx_draws_jitter_y_first(tmp_path):
    """The same statement as host code: g++ evaluates function arguments right to left."""
    gxx = shutil.which("g++")
    inc = "/usr/local/cuda/include"
    if not gxx or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    src = tmp_path / "probe.cpp"
    src.write_text('#include <cuda_runtime.h>\n#include <cstdio>\n#include <cstring>\n' + _header().replace("__host__ __device__ __inline__", "inline") +
                   "int main()\n{\n    float out[2];" + BODY +
                   '    unsigned int b[2];\n    std::memcpy(b, out, 8);\n    std::printf("%u %u\\n", b[0], b[1]);\n    return 0;\n}\n')
    exe = tmp_path / "probe"
    r = subprocess.run([gxx, "-O1", "-I", inc, "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    x, y = (int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split())
    first, second = _draws()
    assert (x, y) == (second, first)
