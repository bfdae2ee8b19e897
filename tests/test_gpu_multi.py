"""pt_group / pt_launch_multi through the C ABI (csrc/multigpu.cu): several GPUs render one image, bit-identical to the same
launches issued one after the other with pt_launch on one GPU.

On a box with ONE GPU the group lists the same device several times (peer transport): plan, class-range launches, slice
exchange, ordered fold and gather all run exactly as between devices, only the copies stay on the device.  With >= 2 GPUs
the same checks run across devices with both transports, from a C++ host (ptb_render --gpus) and with one process per
GPU under torchrun (NCCL)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from acgpathtracing_b200 import scenes as S  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_GPUS = torch.cuda.device_count()


@pytest.fixture(scope="module")
def api():
    from acgpathtracing_b200 import api as A
    return A


def _sequential(api, arr, w, h, spp, md, frames, first=0, acc0=None):
    """The reference order: one pt_launch per frame on one GPU, folding in place."""
    gs = api.Scene(arr)
    try:
        acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0") if acc0 is None else acc0.clone()
        fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
        for f in range(first, first + frames):
            p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
            p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), fb.data_ptr()
            api.launch(p, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return acc, fb
    finally:
        gs.destroy()


def _multi(api, group, gs, w, h, spp, md, frames, first=0, acc0=None, want_fb=True):
    acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0") if acc0 is None else acc0.clone()
    fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda:0")
    p = api.make_params(w, h, spp, md, frame=first, nee=True, importance=True, scene=gs)
    p.accumulationBuffer, p.frameBuffer = acc.data_ptr(), (fb.data_ptr() if want_fb else None)
    group.launch(p, frames, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return acc, fb


def _same(a, b):
    return torch.equal(a.view(torch.int32), b.view(torch.int32))


@pytest.mark.parametrize("n_ranks", [1, 2, 3, 8])
def test_launch_multi_emulated_ranks_equal_sequential_launches(api, n_ranks):
    """Ragged image (not a multiple of the 8x4 tile), every plan shape: fewer frames than ranks (tile shares only), whole
    rounds, leftovers whose sub-shares span two frames; then a second call that resumes at first frame > 0."""
    arr = S.cornell32()
    w, h, spp, md = 203, 117, 3, 12
    torch.cuda.set_device(0)
    group = api.Group.local([0] * n_ranks)
    try:
        assert group.info() == dict(n_ranks=n_ranks, n_local=n_ranks, first_rank=0, transport="peer")
        gs = group.create_scene(arr)
        for frames in (1, 3, n_ranks, 2 * n_ranks + 3):
            want_acc, want_fb = _sequential(api, arr, w, h, spp, md, frames)
            acc, fb = _multi(api, group, gs, w, h, spp, md, frames)
            assert _same(acc, want_acc), (n_ranks, frames)
            assert torch.equal(fb, want_fb), (n_ranks, frames)
        # resume: 5 frames, then 4 more on top of the running average
        a1, _ = _multi(api, group, gs, w, h, spp, md, 5)
        a2, f2 = _multi(api, group, gs, w, h, spp, md, 4, first=5, acc0=a1)
        want_acc, want_fb = _sequential(api, arr, w, h, spp, md, 9)
        assert _same(a2, want_acc) and torch.equal(f2, want_fb)
        t = group.last_timing(0)
        assert t["kernels"] >= 1 and t["total_ms"] > 0
    finally:
        group.destroy()


def test_launch_multi_many_frames_passes_and_material_scene(api):
    """More frames than one pass holds (8 whole frames per rank): the fold continues across passes; the material scene runs
    from global memory (dielectric + conductor paths)."""
    torch.cuda.set_device(0)
    for arr, w, h, spp, md, frames, n in ((S.cornell32(), 96, 54, 1, 8, 41, 2), (S.material_scene(2), 160, 90, 2, 28, 5, 3)):
        group = api.Group.local([0] * n)
        try:
            gs = group.create_scene(arr)
            acc, fb = _multi(api, group, gs, w, h, spp, md, frames)
            want_acc, want_fb = _sequential(api, arr, w, h, spp, md, frames)
            assert _same(acc, want_acc) and torch.equal(fb, want_fb)
        finally:
            group.destroy()


def test_launch_multi_host_buffers_and_errors(api):
    arr = S.cornell32()
    w, h, spp, md = 120, 64, 2, 8
    torch.cuda.set_device(0)
    group = api.Group.local([0, 0, 0])
    try:
        with pytest.raises(api.PtError):                              # no scene yet
            group.launch(api.make_params(w, h, spp, md), 1)
        gs = group.create_scene(arr)
        h_acc = np.zeros((h, w, 4), np.float32)
        h_fb = np.zeros((h, w, 4), np.uint8)
        p = api.make_params(w, h, spp, md, frame=0, nee=True, importance=True, scene=gs)
        p.accumulationBuffer, p.frameBuffer = h_acc.ctypes.data, h_fb.ctypes.data
        up, down = group.launch_host(p, 4)
        assert down == w * h * 20 and up == 168                       # one image down; first call uploads only the params
        p.currentFrameIdx = 4
        up, down = group.launch_host(p, 3)
        assert up == 168 + w * h * 16 and down == w * h * 20          # resume: the previous average goes up, slice by slice
        want_acc, want_fb = _sequential(api, arr, w, h, spp, md, 7)
        assert np.array_equal(h_acc.view(np.uint32), want_acc.cpu().numpy().view(np.uint32))
        assert np.array_equal(h_fb, want_fb.cpu().numpy())
        other = api.Scene(arr)
        try:
            q = api.make_params(w, h, spp, md, scene=other)
            q.accumulationBuffer = h_acc.ctypes.data
            with pytest.raises(api.PtError):                          # a scene that is not the group's
                group.launch_host(q, 1)
        finally:
            other.destroy()
        with pytest.raises(api.PtError):
            group.launch_host(p, 0)
        for i in range(3):                                            # replicas are separate scenes with their own counters
            assert group.replica(i).info()["n_tris"] == arr.n_tri
    finally:
        group.destroy()


def test_launch_multi_counters_sum_to_single_gpu(api):
    arr = S.cornell32()
    w, h, spp, md, frames = 128, 72, 2, 8, 5
    torch.cuda.set_device(0)
    single = api.Scene(arr)
    group = api.Group.local([0, 0])
    try:
        single.enable_counters(True)
        acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0")
        for f in range(frames):
            p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=single)
            p.accumulationBuffer = acc.data_ptr()
            api.launch(p)
        want = single.counters()
        gs = group.create_scene(arr)
        reps = [group.replica(i) for i in range(2)]
        for r in reps:
            r.enable_counters(True)
        _multi(api, group, gs, w, h, spp, md, frames)
        got = [r.counters() for r in reps]
        for key in ("n_samples", "n_radiance", "n_shadow", "n_node", "n_tri", "n_miss"):
            assert got[0][key] + got[1][key] == want[key], key
    finally:
        single.destroy()
        group.destroy()


@pytest.mark.skipif(N_GPUS < 2, reason="needs two GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_launch_multi_two_devices(api, transport):
    arr = S.cornell32()
    w, h, spp, md = 320, 180, 4, 16
    torch.cuda.set_device(0)
    n = min(N_GPUS, 4)
    group = api.Group.local(list(range(n)), api.PT_TRANSPORT_NCCL if transport == "nccl" else api.PT_TRANSPORT_PEER)
    try:
        gs = group.create_scene(arr)
        for frames in (1, n + 1, 2 * n):
            acc, fb = _multi(api, group, gs, w, h, spp, md, frames)
            want_acc, want_fb = _sequential(api, arr, w, h, spp, md, frames)
            assert _same(acc, want_acc) and torch.equal(fb, want_fb), (transport, frames)
    finally:
        group.destroy()


@pytest.mark.skipif(N_GPUS < 2, reason="needs two GPUs")
def test_process_per_gpu_under_torchrun(tmp_path):
    """One process per GPU, NCCL inside the library: tests/multi_rank_check.py compares every rank's image with the
    sequential launches of rank 0."""
    n = min(N_GPUS, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "multi_rank_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "multi_rank_check ok" in r.stdout


def _driver():
    return os.path.join(ROOT, "acgpathtracing_b200", "ptb_render")


def test_cpp_host_renders_on_a_group_bit_identically(api, tmp_path):
    """The C++ host (csrc/ptb_render.cu, the counterpart of the reference's main) with --gpus: pt_group_create_local +
    pt_group_scene_create + pt_launch_multi.  Its accumulation file must equal the single-GPU run's, byte for byte."""
    if not os.path.exists(_driver()):
        pytest.skip("ptb_render not built")
    arr = S.cornell32()
    obj = tmp_path / "cornell32.obj"
    obj.write_text(arr.obj_text)
    (tmp_path / "cornell32.mtl").write_text(arr.mtl_text)
    outs = {}
    devices = {"1": "0", "3": "0,0,0"}
    if N_GPUS >= 2:
        devices["2x"] = "0,1"
    for key, dev in devices.items():
        out = tmp_path / f"acc_{key}.bin"
        cmd = [_driver(), "--obj", str(obj), "--width", "200", "--height", "112", "--spp", "4", "--depth", "16", "--direct",
               "--importance", "--frames", "7", "--out", str(tmp_path / f"img_{key}.ppm"), "--accum-out", str(out)]
        if key != "1":
            cmd += ["--devices", dev]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0, r.stdout
        outs[key] = (out.read_bytes(), (tmp_path / f"img_{key}.ppm").read_bytes())
    for key in outs:
        assert outs[key] == outs["1"], key
