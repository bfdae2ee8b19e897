"""GPU parity tests: the CUDA path through the C-ABI (include/ptb200.h) against the CPU oracle.

Bars (SURVEY 8c / north star):
  * bit-exact: RNG streams (pixel seed after jitter, final path / pixel seeds), Morton keys, sort permutation,
    LBVH topology + node boxes, primary-hit (primitive id, t bits), any-hit bits, work counters;
  * radiance: the GPU kernels use the same portable transcendentals as the oracle's "det" mode
    (csrc/pt_detmath.h) and are compiled without fused multiply-add contraction, so accumulation buffers are
    required to be BIT-IDENTICAL too (RADIANCE_RTOL = 0); the north star's looser bound (1e-4 relative at 1 spp)
    is asserted against the oracle's glibc mode, where sinf/cosf/acosf differ in the last place.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_api as O
from acgpathtracing_b200 import scenes as S
from acgpathtracing_b200.types import PathTraceParams

pytestmark = pytest.mark.gpu

RADIANCE_RTOL = 0.0          # vs the oracle's det-math mode: bit-exact
LIBM_RTOL = 1e-4             # vs the oracle's glibc mode (reference source semantics), per channel
LIBM_OUTLIER_BUDGET = 1e-4   # fraction of pixels allowed outside LIBM_RTOL (attributed pixel by pixel in tests/test_tolerance.py)


@pytest.fixture(scope="module")
def api():
    from acgpathtracing_b200 import api as A
    return A


def _scene(name):
    return {"cornell": S.cornell32, "material": lambda: S.material_scene(2), "tess40": lambda: S.tessellated_scene(40),
            "tess300": lambda: S.tessellated_scene(300)}[name]()


@pytest.fixture(scope="module")
def scenes_cache():
    return {}


def _get(scenes_cache, api, port, name):
    if name not in scenes_cache:
        arr = _scene(name)
        scenes_cache[name] = (arr, api.Scene(arr), O.OracleScene(port, arr))
    return scenes_cache[name]


# ---- LBVH: Morton keys, radix sort, Karras tree, refit ----------------------------------------------------------
@pytest.mark.parametrize("name", ["cornell", "material", "tess40", "tess300"])
def test_lbvh_bit_exact(api, port, scenes_cache, name):
    arr, gs, osc = _get(scenes_cache, api, port, name)
    g, o = gs.bvh_export(), osc.bvh()
    assert np.array_equal(g["bounds"].view(np.uint32), o["bounds"].view(np.uint32))
    assert np.array_equal(g["morton"], o["morton"])
    assert np.array_equal(g["sorted_keys"], o["sorted_keys"])
    assert np.array_equal(g["sorted_prim"], o["sorted_prim"])
    assert np.array_equal(g["parent"], o["parent"])
    gn, on = g["nodes"].copy(), o["nodes"].copy()
    gn[:, [11, 15]] = 0   # padding words
    on[:, [11, 15]] = 0
    assert np.array_equal(gn, on)
    gt, ot = g["tris"].copy(), o["tris"].copy()
    gt[:, 11] = 0
    ot[:, 11] = 0
    assert np.array_equal(gt, ot)


def _rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    lo = scene.verts[:, :3].min(0) - 50
    hi = scene.verts[:, :3].max(0) + 50
    org = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    d[: n // 20] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 20)] * rng.choice([-1, 1], (n // 20, 1))
    tri = scene.indices[rng.integers(0, scene.n_tri, n // 4)]
    bary = rng.dirichlet((1, 1, 1), n // 4)
    p = (scene.verts[tri[:, 0], :3] * bary[:, :1] + scene.verts[tri[:, 1], :3] * bary[:, 1:2] +
         scene.verts[tri[:, 2], :3] * bary[:, 2:3])
    org[n // 2: n // 2 + n // 4] = p.astype(np.float32)
    return org, d


@pytest.mark.parametrize("name", ["cornell", "material", "tess40"])
def test_traversal_bit_exact(api, port, scenes_cache, name):
    arr, gs, osc = _get(scenes_cache, api, port, name)
    n = 200_000
    org, d = _rays(arr, n, 11)
    gp, gt, ga, gc = gs.trace(org, d, 0.01, 1e16, want_any=True)
    op, ot, oa, oc = osc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
    assert np.array_equal(gp, op)
    assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32))
    assert np.array_equal(ga, oa)
    assert gc.tolist() == oc.tolist()                      # node fetches and triangle tests, ray by ray order-free sums
    # brute force over all triangles agrees as well (BVH culling is conservative)
    bp, bt, ba, _ = osc.trace(org[:50_000], d[:50_000], 0.01, 1e16, use_bvh=False, want_any=True)
    assert np.array_equal(gp[:50_000], bp) and np.array_equal(gt[:50_000].view(np.uint32), bt.view(np.uint32))
    assert np.array_equal(ga[:50_000], ba)
    tmax = np.random.default_rng(5).uniform(1, 600, n).astype(np.float32)
    _, _, ga, _ = gs.trace(org, d, 0.01, tmax, want_any=True)
    _, _, oa, _ = osc.trace(org, d, 0.01, tmax, use_bvh=True, want_any=True)
    assert np.array_equal(ga, oa)


# ---- RNG streams + primary hits -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name,w,h", [("cornell", 512, 512), ("material", 320, 180), ("cornell", 253, 131)])
def test_rng_streams_and_primary_hits(api, port, scenes_cache, name, w, h):
    arr, gs, osc = _get(scenes_cache, api, port, name)
    p = api.make_params(w, h, 3, 5, frame=2, nee=True, importance=True, scene=gs)
    g = gs.debug_primary(p)
    gf = gs.debug_final_seeds(p)
    o = osc.render(p, math_mode=1, use_bvh=True, probes=True)
    assert np.array_equal(g["seed_after_jitter"], o["seed_after_jitter"])
    assert np.array_equal(g["primary_dir"].view(np.uint32), o["primary_dir"].view(np.uint32))
    assert np.array_equal(g["prim"], o["prim"])
    assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    assert np.array_equal(gf["final_path_seed"], o["final_path_seed"])
    assert np.array_equal(gf["final_pixel_seed"], o["final_pixel_seed"])
    assert (g["prim"] >= 0).mean() > 0.2


def test_rng_known_answers(api, port, scenes_cache):
    """tea<4> / LCG known answers from the reference's own cuda/random.h through the kernel."""
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    p = api.make_params(512, 512, 1, 3, frame=0, scene=gs)
    g = gs.debug_primary(p)
    s = 1576399551                                # tea<4>(0, 0)
    for _ in range(2):
        s = (1664525 * s + 1013904223) & 0xFFFFFFFF
    assert int(g["seed_after_jitter"][0, 0]) == s == 4105309513
    p3 = api.make_params(512, 512, 1, 3, frame=3, scene=gs)
    g3 = gs.debug_primary(p3)
    s = 2391278879                                # tea<4>(262143, 3)
    for _ in range(2):
        s = (1664525 * s + 1013904223) & 0xFFFFFFFF
    assert int(g3["seed_after_jitter"][511, 511]) == s


# ---- radiance ---------------------------------------------------------------------------------------------------------
RENDER_CASES = [
    # scene, w, h, spp, maxDepth, NEE, IS, launches
    ("cornell", 512, 512, 1, 3, False, True, 1),      # BASELINE config C1
    ("cornell", 480, 270, 8, 16, True, True, 3),      # C2 shape (16 bounces, IS + NEE), reduced size
    ("material", 320, 180, 8, 28, True, True, 2),     # C3 shape
    ("material", 200, 120, 4, 6, False, False, 1),    # uniform hemisphere sampling
    ("tess40", 240, 136, 4, 8, True, True, 1),        # C4 shape, reduced
    ("cornell", 97, 61, 5, 1, True, True, 2),         # ragged tiles, minimum depth
]


@pytest.mark.parametrize("case", RENDER_CASES, ids=lambda c: "-".join(str(x) for x in c))
def test_render_matches_oracle(api, port, scenes_cache, case):
    name, w, h, spp, md, nee, imp, frames = case
    arr, gs, osc = _get(scenes_cache, api, port, name)
    gs.enable_counters(True)
    gs.counters(reset=True)
    acc_g = acc_o = None
    tot = None
    try:
        for f in range(frames):
            p = api.make_params(w, h, spp, md, frame=f, nee=nee, importance=imp, scene=gs)
            g = gs.launch_host(p, accum=acc_g)
            o = osc.render(p, accum=acc_o, math_mode=1, use_bvh=True)
            acc_g, acc_o = g["accum"], o["accum"]
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), f"frame {f}"
            assert np.array_equal(g["fb"], o["fb"]), f"frame {f}"
            tot = o["counters"] if tot is None else {k: (tot[k] + v if not isinstance(v, list) else
                                                         [a + b for a, b in zip(tot[k], v)]) for k, v in o["counters"].items()}
        assert gs.counters(reset=True) == tot
    finally:
        gs.enable_counters(False)


def test_uncounted_kernel_equals_counted(api, port, scenes_cache):
    arr, gs, osc = _get(scenes_cache, api, port, "material")
    p = api.make_params(192, 108, 6, 12, frame=1, nee=True, scene=gs)
    prev = np.random.default_rng(0).random((108, 192, 4), dtype=np.float32)
    a = gs.launch_host(p, accum=prev.copy())
    gs.enable_counters(True)
    b = gs.launch_host(p, accum=prev.copy())
    gs.enable_counters(False)
    gs.counters(reset=True)
    assert np.array_equal(a["accum"].view(np.uint32), b["accum"].view(np.uint32))
    assert np.array_equal(a["fb"], b["fb"])


@pytest.mark.parametrize("name,w,h,md,nee", [("cornell", 512, 512, 3, False), ("cornell", 960, 540, 16, True), ("material", 960, 540, 28, True)],
                         ids=["c1", "c2-shape", "c3-shape"])
def test_radiance_vs_reference_libm_semantics(api, port, scenes_cache, name, w, h, md, nee, capsys):
    """North-star tolerance: 1e-4 relative at 1 spp with identical seeds against the reference-source semantics
    (glibc sinf/cosf/acosf/powf), on the shapes of configs 1-3.  The measured outlier fraction is printed and must stay within
    the survey's budget of 1e-4 of the pixels; tests/test_tolerance.py attributes every outlier (a flipped Russian-roulette /
    Fresnel / shadow decision, or an ill-conditioned grazing term) on the CPU, where both math modes are available."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    p = api.make_params(w, h, 1, md, frame=0, nee=nee, importance=True, scene=gs)
    g = gs.launch_host(p)
    o = osc.render(p, math_mode=0, use_bvh=True)
    ga, oa = g["accum"][..., :3], o["accum"][..., :3]
    bad = (np.abs(ga - oa) > LIBM_RTOL * np.maximum(np.abs(oa), 1e-3)).any(axis=-1)
    with capsys.disabled():
        print(f"\n[tolerance, GPU vs reference semantics] {name} {w}x{h} depth {md}: {int(bad.sum())} of {bad.size} pixels outside "
              f"{LIBM_RTOL:g} relative ({bad.mean():.2e}; budget {LIBM_OUTLIER_BUDGET:g})")
    assert bad.mean() <= LIBM_OUTLIER_BUDGET
    assert np.abs(g["fb"].astype(int) - o["fb"].astype(int))[~bad].max() <= 1      # framebuffer bytes: +-1 LSB


# ---- edge cases ---------------------------------------------------------------------------------------------------------
def test_empty_scene_and_background(api, port):
    empty = S.SceneArrays(np.ones((3, 4), np.float32), np.zeros((0, 3), np.uint32), np.zeros(0, np.uint32),
                          [dict(name="m", Kd=(1, 1, 1), Ke=(0, 0, 0))])
    gs = api.Scene(empty, background=(0.25, 0.5, 0.75))
    p = api.make_params(2, 2, 2, 3, scene=gs)
    g = gs.launch_host(p)
    assert np.array_equal(g["accum"], np.broadcast_to(np.array([0.25, 0.5, 0.75, 1.0], np.float32), (2, 2, 4)))
    assert np.array_equal(g["fb"], np.broadcast_to(np.array([137, 188, 225, 255], np.uint8), (2, 2, 4)))


def test_single_triangle_scene(api, port):
    one = S.SceneArrays(np.array([[100, 100, 300, 1], [500, 100, 300, 1], [300, 500, 300, 1]], np.float32),
                        np.array([[0, 1, 2]], np.uint32), np.zeros(1, np.uint32),
                        [dict(name="m", Kd=(0.5, 0.6, 0.7), Ke=(0, 0, 0))])
    gs = api.Scene(one, background=(0.1, 0.2, 0.3))
    osc = O.OracleScene(port, one, background=(0.1, 0.2, 0.3))
    p = api.make_params(64, 64, 4, 4, nee=True, scene=gs)
    g = gs.launch_host(p)
    o = osc.render(p, math_mode=1, use_bvh=True)
    assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32))
    prim, t, anyh, _ = gs.trace(np.array([[300, 200, -1]], np.float32), np.array([[0, 0, 1]], np.float32), 0.01, 1e16, True)
    assert prim.tolist() == [0] and t[0] == 301.0 and anyh.tolist() == [1]


def test_max_depth_28_and_many_frames(api, port, scenes_cache):
    arr, gs, osc = _get(scenes_cache, api, port, "material")
    acc_g = acc_o = None
    for f in range(5):
        p = api.make_params(96, 54, 2, 28, frame=f, nee=True, scene=gs)
        acc_g = gs.launch_host(p, accum=acc_g)["accum"]
        acc_o = osc.render(p, accum=acc_o, math_mode=1, use_bvh=True)["accum"]
    assert np.array_equal(acc_g.view(np.uint32), acc_o.view(np.uint32))


def test_error_codes(api, port, scenes_cache):
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    p = api.make_params(16, 16, 1, 1, scene=gs)
    p.handle = 12345
    p.accumulationBuffer = 1
    with pytest.raises(api.PtError) as e:
        api.launch(p)
    assert e.value.code == api.PT_ERR_INVALID
    bad = S.cornell32()
    bad.mat_ids = bad.mat_ids.copy()
    bad.mat_ids[3] = 0xFFFFFFFF          # tinyobj's "no usemtl" id
    with pytest.raises(api.PtError) as e:
        api.Scene(bad)
    assert e.value.code == api.PT_ERR_INVALID
    bad = S.cornell32()
    bad.indices = bad.indices.copy()
    bad.indices[0, 0] = 10_000
    with pytest.raises(api.PtError):
        api.Scene(bad)
    p = api.make_params(16, 16, 0, 1, scene=gs)
    with pytest.raises(api.PtError):
        gs.launch_host(p)


# ---- device-buffer launches, partition, determinism at BASELINE size ----------------------------------------------
def _dev_buffers(w, h):
    import torch
    acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    return acc, fb


def _launch_dev(api, gs, p, acc, fb, rank=None, n=None):
    import torch
    q = PathTraceParams.from_buffer_copy(p)
    q.handle = gs.handle
    q.accumulationBuffer = acc.data_ptr()
    q.frameBuffer = fb.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    if rank is None:
        api.launch(q, stream)
    else:
        api.launch_partition(q, rank, n, stream)


@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_partition_union_is_bit_identical(api, port, scenes_cache, n_ranks):
    import torch
    from acgpathtracing_b200 import partition as P
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    w, h = 250, 122
    p = api.make_params(w, h, 4, 16, frame=0, nee=True, scene=gs)
    full_acc, full_fb = _dev_buffers(w, h)
    _launch_dev(api, gs, p, full_acc, full_fb)
    cap = api.partition_capacity(w, h, n_ranks)
    assert cap == P.capacity(w, h, n_ranks)
    comp_acc = torch.zeros((n_ranks, cap, 4), dtype=torch.float32, device="cuda")
    comp_fb = torch.zeros((n_ranks, cap, 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    owner = torch.from_numpy(P.owner_map(w, h, n_ranks)).cuda()
    for r in range(n_ranks):
        acc, fb = _dev_buffers(w, h)
        acc.fill_(-7.0)
        _launch_dev(api, gs, p, acc, fb, r, n_ranks)
        torch.cuda.synchronize()
        mine = owner == r
        assert torch.equal(acc[mine].view(torch.int32), full_acc[mine].view(torch.int32))     # owned pixels rendered
        assert bool((acc[~mine] == -7.0).all())                                               # others untouched
        api.partition_pack(acc.data_ptr(), fb.data_ptr(), comp_acc[r].data_ptr(), comp_fb[r].data_ptr(), w, h, r, n_ranks, stream)
        torch.cuda.synchronize()
        assert np.array_equal(comp_acc[r].cpu().numpy(), P.pack_numpy(acc.cpu().numpy(), w, h, r, n_ranks))
    out_acc, out_fb = _dev_buffers(w, h)
    api.partition_unpack(comp_acc.data_ptr(), comp_fb.data_ptr(), out_acc.data_ptr(), out_fb.data_ptr(), w, h, n_ranks, stream)
    torch.cuda.synchronize()
    assert torch.equal(out_acc.view(torch.int32), full_acc.view(torch.int32))
    assert torch.equal(out_fb, full_fb)
    assert np.array_equal(P.unpack_numpy(comp_acc.cpu().numpy(), w, h, n_ranks), full_acc.cpu().numpy())


def test_full_size_c2_properties(api, port, scenes_cache):
    """BASELINE config C2 shape at full resolution (1920x1080, 16 bounces, IS + NEE), one 128-spp launch:
    determinism (two launches bit-identical), a crop of the image bit-identical to the oracle (pixels are
    independent), every pixel written, and the running-average fold of a second launch."""
    import torch
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    w, h, spp = 1920, 1080, 128
    p = api.make_params(w, h, spp, 16, frame=0, nee=True, scene=gs)
    a1, f1 = _dev_buffers(w, h)
    a2, f2 = _dev_buffers(w, h)
    a1.fill_(float("nan"))
    _launch_dev(api, gs, p, a1, f1)
    _launch_dev(api, gs, p, a2, f2)
    torch.cuda.synchronize()
    assert torch.equal(a1.view(torch.int32), a2.view(torch.int32)) and torch.equal(f1, f2)
    assert bool(torch.isfinite(a1).all()) and bool((a1[..., 3] == 1.0).all()) and bool((f1[..., 3] == 255).all())
    x0, x1, y0, y1 = 900, 932, 500, 516
    o = osc.render(p, math_mode=1, use_bvh=True, region=(x0, x1, y0, y1))
    ga = a1[y0:y1, x0:x1].cpu().numpy()
    assert np.array_equal(ga.view(np.uint32), o["accum"][y0:y1, x0:x1].view(np.uint32))
    assert np.array_equal(f1[y0:y1, x0:x1].cpu().numpy(), o["fb"][y0:y1, x0:x1])
    # second launch folds into the running average with weight 1/2 (pathTracerPrograms.cu:803-810)
    p1 = api.make_params(w, h, spp, 16, frame=1, nee=True, scene=gs)
    _launch_dev(api, gs, p1, a1, f1)
    torch.cuda.synchronize()
    o1 = osc.render(p1, accum=o["accum"], math_mode=1, use_bvh=True, region=(x0, x1, y0, y1))
    assert np.array_equal(a1[y0:y1, x0:x1].cpu().numpy().view(np.uint32), o1["accum"][y0:y1, x0:x1].view(np.uint32))
    mean = a1[..., :3].mean().item()
    assert 0.05 < mean < 2.0


# ---- headless driver (the reference's main() on top of the C ABI) + image output + checkpoint / resume ---------------
def _read_ppm(path):
    raw = open(path, "rb").read()
    parts = raw.split(b"\n", 3)
    assert parts[0] == b"P6" and parts[2] == b"255"
    w, h = (int(x) for x in parts[1].split())
    return np.frombuffer(parts[3], np.uint8).reshape(h, w, 3)


def test_headless_driver_matches_oracle_and_resumes(api, port, scenes_cache, tmp_path):
    import os
    import subprocess
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "acgpathtracing_b200", "ptb_render")
    obj = tmp_path / "cornell32.obj"
    obj.write_text(arr.obj_text)
    (tmp_path / "cornell32.mtl").write_text(arr.mtl_text)
    w, h, spp, md = 96, 54, 4, 16
    base = [exe, "--obj", str(obj), "--width", str(w), "--height", str(h), "--spp", str(spp), "--depth", str(md),
            "--direct", "--importance"]
    r = subprocess.run(base + ["--frames", "2", "--out", str(tmp_path / "two.ppm"), "--accum-out", str(tmp_path / "two.acc")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Frame Render Time:" in r.stdout and "Total Samples 8" in r.stdout        # PathTracerMain.cpp:726,738
    acc_o = None
    for f in range(2):
        p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
        o = osc.render(p, accum=acc_o, math_mode=1, use_bvh=True)
        acc_o = o["accum"]
    assert np.array_equal(_read_ppm(tmp_path / "two.ppm"), o["fb"][::-1, :, :3])     # saveImage flips rows
    # one frame, checkpoint, resume for one more frame == two frames in one go
    r1 = subprocess.run(base + ["--frames", "1", "--out", str(tmp_path / "a.ppm"), "--accum-out", str(tmp_path / "a.acc")],
                        capture_output=True, text=True, timeout=300)
    r2 = subprocess.run(base + ["--frames", "1", "--out", str(tmp_path / "b.ppm"), "--accum-in", str(tmp_path / "a.acc"),
                                "--accum-out", str(tmp_path / "b.acc")], capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    assert "Resuming at frame 1" in r2.stdout
    assert open(tmp_path / "b.acc", "rb").read() == open(tmp_path / "two.acc", "rb").read()
    assert open(tmp_path / "b.ppm", "rb").read() == open(tmp_path / "two.ppm", "rb").read()
    bad = subprocess.run([exe, "--obj", str(tmp_path / "missing.obj")], capture_output=True, text=True, timeout=120)
    assert bad.returncode == 1 and "Caught exception:" in bad.stderr                # main's catch block, :742-744


# ---- BASELINE configs C3 / C4 / C5 at full size: size-independent properties + crops against the oracle ---------------
def _crop_check(api, gs, osc, w, h, spp, md, crop, frames=1, pipelines=(0,)):
    import torch
    x0, x1, y0, y1 = crop
    acc_o = None
    outs = []
    for pl in pipelines:
        gs.set_pipeline(pl)
        acc, fb = _dev_buffers(w, h)
        for f in range(frames):
            p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
            _launch_dev(api, gs, p, acc, fb)
        torch.cuda.synchronize()
        outs.append((acc.clone(), fb.clone()))
    gs.set_pipeline(0)
    for f in range(frames):
        p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
        o = osc.render(p, accum=acc_o, math_mode=1, use_bvh=True, region=crop)
        acc_o = o["accum"]
    for acc, fb in outs:
        assert bool(torch.isfinite(acc).all()) and bool((acc[..., 3] == 1.0).all())
        ga = acc[y0:y1, x0:x1].cpu().numpy()
        assert np.array_equal(ga.view(np.uint32), o["accum"][y0:y1, x0:x1].view(np.uint32))
        assert np.array_equal(fb[y0:y1, x0:x1].cpu().numpy(), o["fb"][y0:y1, x0:x1])
    for acc, fb in outs[1:]:                       # both render kernels agree on every pixel
        assert torch.equal(acc.view(torch.int32), outs[0][0].view(torch.int32)) and torch.equal(fb, outs[0][1])
    return outs[0]


def test_full_size_c3_material_scene(api, port):
    """C3 shape: dielectric + conductor spheres in the Cornell box, 1920x1080, 28 bounces, IS + NEE; one launch at a
    reduced 16 spp (the per-pixel arithmetic does not depend on the image size or on the other pixels).  A crop through
    the glass sphere and one through the metal sphere are bit-identical to the oracle; wavefront == fallback kernel."""
    arr = S.material_scene(3)
    gs, osc = api.Scene(arr), O.OracleScene(port, arr)
    try:
        _crop_check(api, gs, osc, 1920, 1080, 16, 28, (700, 724, 300, 312), frames=2, pipelines=(0, 1))
        _crop_check(api, gs, osc, 1920, 1080, 16, 28, (1180, 1204, 330, 342))
    finally:
        gs.destroy()


def test_full_size_c4_two_million_triangles(api, port):
    """C4: the 2.0 M-triangle tessellated scene.  LBVH products bit-identical to the CPU statement (Morton keys, sort
    permutation, topology, node boxes), 4096 random rays through the device traversal == CPU LBVH == brute force on a
    subset, and a 1080p 8-bounce crop bit-identical to the oracle."""
    arr = S.tessellated_scene(1000)
    assert arr.n_tri >= 2_000_000
    gs, osc = api.Scene(arr), O.OracleScene(port, arr)
    try:
        g, o = gs.bvh_export(), osc.bvh()
        assert np.array_equal(g["sorted_keys"], o["sorted_keys"])
        assert np.all(np.diff(g["sorted_keys"].astype(np.uint64)) > 0)            # strictly ascending: keys are unique
        assert np.array_equal(np.sort(g["sorted_prim"]), np.arange(arr.n_tri, dtype=g["sorted_prim"].dtype))
        assert np.array_equal(g["parent"], o["parent"])
        gn, on = g["nodes"].copy(), o["nodes"].copy()
        gn[:, [11, 15]] = 0
        on[:, [11, 15]] = 0
        assert np.array_equal(gn, on)
        gpl, opl = gs.bvh_export_ploc(), osc.bvh_ploc()                           # the tree both sides walk by default
        gn, on = gpl["nodes"].copy(), opl["nodes"].copy()
        gn[:, [11, 15]] = 0
        on[:, [11, 15]] = 0
        assert gpl["depth"] == opl["depth"] and np.array_equal(gn, on)
        assert gs.tree_in_use() == osc.set_tree(0) == "ploc"
        org, d = _rays(arr, 4096, 23)
        gp, gt, ga, gc = gs.trace(org, d, 0.01, 1e16, want_any=True)
        op, ot, oa, oc = osc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
        assert np.array_equal(gp, op) and np.array_equal(gt.view(np.uint32), ot.view(np.uint32)) and np.array_equal(ga, oa)
        assert gc.tolist() == oc.tolist()
        bp, bt, ba, _ = osc.trace(org[:96], d[:96], 0.01, 1e16, use_bvh=False, want_any=True)
        assert np.array_equal(gp[:96], bp) and np.array_equal(gt[:96].view(np.uint32), bt.view(np.uint32))
        _crop_check(api, gs, osc, 1920, 1080, 4, 8, (940, 964, 420, 432), pipelines=(0, 1))
        assert gs.last_launch_info()["traversal"] == "binary"                     # pipeline 1 ran last
        import torch
        acc_c, _ = _crop_check(api, gs, osc, 1920, 1080, 4, 8, (940, 964, 420, 432))
        assert gs.last_launch_info()["traversal"] == "wide4-quant"                # auto: the records do not fit in L2
        for trav, expect in ((2, "binary"), (3, "binary-quant"), (1, "binary")):
            gs.set_traversal(trav)
            acc_t, _ = _crop_check(api, gs, osc, 1920, 1080, 4, 8, (940, 964, 420, 432))
            assert gs.last_launch_info()["traversal"] == expect
            assert torch.equal(acc_t.view(torch.int32), acc_c.view(torch.int32))  # every pixel of the 1080p frame
    finally:
        gs.destroy()


def test_full_size_c5_4k(api, port, scenes_cache):
    """C5 shape: 3840x2160, 28 bounces, IS + NEE, one 32-spp launch + a second folded in: determinism, every pixel
    written, crop bit-identical to the oracle, and the 8-rank pixel partition reproduces the image."""
    import torch
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    w, h, spp, md = 3840, 2160, 32, 28
    acc, fb = _crop_check(api, gs, osc, w, h, spp, md, (1900, 1932, 1000, 1016), frames=2)
    acc2, fb2 = _dev_buffers(w, h)
    for f in range(2):
        p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
        for r in range(8):
            _launch_dev(api, gs, p, acc2, fb2, rank=r, n=8)
    torch.cuda.synchronize()
    assert torch.equal(acc.view(torch.int32), acc2.view(torch.int32)) and torch.equal(fb, fb2)


# ---- frame-parallel rendering: raw per-launch means + ordered fold == launches one after the other -------------------
@pytest.mark.parametrize("name,w,h,spp,md,frames", [("cornell", 250, 122, 3, 16, 5), ("material", 160, 90, 2, 28, 3)])
def test_raw_means_plus_fold_equal_sequential_launches(api, port, scenes_cache, name, w, h, spp, md, frames):
    import torch
    from acgpathtracing_b200 import multirank as M
    arr, gs, osc = _get(scenes_cache, api, port, name)
    stream = torch.cuda.current_stream().cuda_stream
    acc, fb = _dev_buffers(w, h)
    for f in range(frames):
        _launch_dev(api, gs, api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs), acc, fb)
    means = torch.zeros((frames, h, w, 4), dtype=torch.float32, device="cuda")
    junk = torch.full((h, w, 4), 7, dtype=torch.uint8, device="cuda")
    for f in range(frames):
        p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
        q = PathTraceParams.from_buffer_copy(p)
        q.handle, q.accumulationBuffer, q.frameBuffer = gs.handle, means[f].data_ptr(), junk.data_ptr()
        if f % 2 == 0:
            api.launch_raw(q, 0, 1, stream)
        else:                                           # the tile partition composes with raw launches
            for r in range(3):
                api.launch_raw(q, r, 3, stream)
    assert bool((junk == 7).all())                      # raw launches leave the frame buffer alone
    acc2, fb2 = _dev_buffers(w, h)
    api.fold_frames(acc2.data_ptr(), fb2.data_ptr(), [means[f].data_ptr() for f in range(frames)], 0, w, h, stream)
    torch.cuda.synchronize()
    assert torch.equal(acc.view(torch.int32), acc2.view(torch.int32)) and torch.equal(fb, fb2)
    # resume: fold the first two, then the rest on top of the stored running average
    acc3, fb3 = _dev_buffers(w, h)
    api.fold_frames(acc3.data_ptr(), None, [means[f].data_ptr() for f in range(2)], 0, w, h, stream)
    api.fold_frames(acc3.data_ptr(), fb3.data_ptr(), [means[f].data_ptr() for f in range(2, frames)], 2, w, h, stream)
    torch.cuda.synchronize()
    assert torch.equal(acc.view(torch.int32), acc3.view(torch.int32)) and torch.equal(fb, fb3)
    # the CPU mirror used by the gloo tests computes the same bits
    cpu = M.fold_numpy([means[f].cpu().numpy() for f in range(frames)])
    assert np.array_equal(cpu.view(np.uint32), acc.cpu().numpy().view(np.uint32))
    with pytest.raises(api.PtError):
        api.fold_frames(acc3.data_ptr(), None, [means[0].data_ptr()] * 65, 0, w, h, stream)


@pytest.mark.parametrize("name,md", [("cornell", 16), ("material", 28)], ids=["c2-shape", "c3-shape"])
def test_converged_image_vs_reference_semantics(api, port, scenes_cache, name, md, capsys):
    """North-star bound for converged images (config 3 is quoted at 4096 spp): 32 launches x 128 spp, running average, GPU (portable
    transcendentals) against the oracle in glibc mode = the reference source's semantics.  Individual paths may differ
    after a flipped discrete decision, the converged image may not: RMSE of linear radiance <= 1e-3 x mean luminance,
    frame buffers equal up to 1 LSB in at most 1 % of the bytes.  Against the oracle's own deterministic mode the image
    is bit-identical (checked on the first 4 launches to keep the CPU side short)."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    w, h, spp, launches = 96, 54, 128, 32
    acc_g = acc_l = acc_d = None
    for f in range(launches):
        p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
        g = gs.launch_host(p, accum=acc_g)
        acc_g = g["accum"]
        o = osc.render(p, accum=acc_l, math_mode=0, use_bvh=True)
        acc_l = o["accum"]
        if f < 4:
            d = osc.render(p, accum=acc_d, math_mode=1, use_bvh=True)
            acc_d = d["accum"]
            assert np.array_equal(acc_g.view(np.uint32), acc_d.view(np.uint32)), f"launch {f}"
    lum = (acc_l[..., :3] * np.array([0.2126, 0.7152, 0.0722], np.float32)).sum(-1).mean()
    rmse = float(np.sqrt(np.mean((acc_g[..., :3].astype(np.float64) - acc_l[..., :3].astype(np.float64)) ** 2)))
    assert rmse <= 1e-3 * lum, (rmse, lum)
    diff = np.abs(g["fb"].astype(np.int32) - o["fb"].astype(np.int32))
    with capsys.disabled():
        print(f"\n[converged, 4096 spp] {name}: RMSE {rmse:.3e} = {rmse / lum:.2e} x mean luminance {lum:.4f}; frame-buffer bytes: "
              f"{(diff > 0).mean():.2%} differ, {(diff > 1).mean():.3%} by more than 1 LSB, max {int(diff.max())}")
    if name == "cornell":
        assert diff.max() <= 1 and (diff > 0).mean() <= 0.01
    else:
        # the glass sphere's caustic paths are rare and bright: at 4096 spp a pixel that took one of them in only one of the two
        # runs still differs visibly after the sRGB curve; the bound on such bytes is statistical, the RMSE bound above is not
        assert (diff > 1).mean() <= 0.01


# ---- other cameras: candidate lists, tile order and tail mode must not depend on the reference's hard-coded view -----
def _with_camera(api, p, eye, lookat, up, fovy):
    from acgpathtracing_b200.types import Float3
    U, V, W = api.camera_uvw_frame(eye, lookat, up, fovy, np.float32(p.width) / np.float32(p.height))
    p.cameraEye = Float3(*eye)
    p.cameraU, p.cameraV, p.cameraW = Float3(*U), Float3(*V), Float3(*W)
    return p


CAMERAS = [
    # eye, lookat, up, fovY, width, height
    ((278.0, 273.0, 100.0), (400.0, 0.0, 559.0), (0.0, 1.0, 0.0), 70.0, 160, 90),     # inside the box, at a floor corner
    ((50.0, 500.0, 50.0), (368.0, 165.0, 351.0), (0.0, 1.0, 0.0), 40.0, 120, 120),    # from under the ceiling at the blocks
    ((278.0, 273.0, -900.0), (278.0, 273.0, 330.0), (0.3, 1.0, 0.0), 35.0, 31, 17),   # rolled, tiny image: huge pixels
    ((900.0, 900.0, -600.0), (278.0, 273.0, 280.0), (0.0, 1.0, 0.0), 25.0, 200, 64),  # from outside, above
    ((278.0, 273.0, 279.0), (278.0, 548.0, 279.5), (0.0, 0.0, 1.0), 100.0, 64, 64),   # looking straight up into the light
    ((278.0, 1.0, 1.0), (278.0, 1.5, 600.0), (0.0, 1.0, 0.0), 120.0, 96, 40),         # grazing along the floor, very wide
]


@pytest.mark.parametrize("cam", CAMERAS, ids=[f"cam{i}" for i in range(len(CAMERAS))])
def test_render_matches_oracle_for_other_cameras(api, port, scenes_cache, cam):
    eye, lookat, up, fovy, w, h = cam
    for name in ("cornell", "material"):
        arr, gs, osc = _get(scenes_cache, api, port, name)
        acc_g = acc_o = None
        for f in range(2):
            p = _with_camera(api, api.make_params(w, h, 6, 12, frame=f, nee=True, importance=True, scene=gs), eye, lookat, up, fovy)
            g = gs.launch_host(p, accum=acc_g)
            o = osc.render(p, accum=acc_o, math_mode=1, use_bvh=True)
            acc_g, acc_o = g["accum"], o["accum"]
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), (name, f)
            assert np.array_equal(g["fb"], o["fb"]), (name, f)
        pp = _with_camera(api, api.make_params(w, h, 2, 3, frame=1, scene=gs), eye, lookat, up, fovy)
        gprim, oprobe = gs.debug_primary(pp), osc.render(pp, math_mode=1, use_bvh=True, probes=True)
        assert np.array_equal(gprim["prim"], oprobe["prim"])
        assert np.array_equal(gprim["t"].view(np.uint32), oprobe["t"].view(np.uint32))


# ---- random triangle soups x random cameras: the candidate lists must stay a superset of what a camera ray can hit -----
def _random_soup(rng, n_tri):
    """Triangles of very different shapes inside the Cornell box volume: big, tiny, slivers, edge-on, shared vertices."""
    verts, faces = [], []
    for k in range(n_tri):
        c = rng.uniform([0, 0, 0], [556, 548, 559])
        kind = k % 5
        if kind == 0:                                   # big
            e1, e2 = rng.normal(size=3) * 250, rng.normal(size=3) * 250
        elif kind == 1:                                 # tiny
            e1, e2 = rng.normal(size=3) * 2, rng.normal(size=3) * 2
        elif kind == 2:                                 # sliver
            e1 = rng.normal(size=3) * 300
            e2 = e1 * rng.uniform(0.2, 0.9) + rng.normal(size=3) * 0.5
        elif kind == 3:                                 # axis-aligned quad half, like the box walls
            a = rng.integers(0, 3)
            e1, e2 = np.zeros(3), np.zeros(3)
            e1[(a + 1) % 3], e2[(a + 2) % 3] = rng.uniform(50, 400), rng.uniform(50, 400)
        else:                                           # medium
            e1, e2 = rng.normal(size=3) * 60, rng.normal(size=3) * 60
        base = len(verts)
        verts += [c, c + e1, c + e2]
        faces.append((base, base + 1, base + 2))
    v = np.ones((len(verts), 4), np.float32)
    v[:, :3] = np.asarray(verts, np.float32)
    mats = [dict(name="white", Kd=(0.8, 0.8, 0.8), Ke=(0, 0, 0)), dict(name="glow", Kd=(0.7, 0.6, 0.5), Ke=(4, 3, 2)),
            dict(name="MetallicThing", Kd=(0.9, 0.8, 0.7), Ke=(0, 0, 0)), dict(name="RefractiveThing", Kd=(1, 1, 1), Ke=(0, 0, 0), Ni=1.5)]
    mat_ids = rng.integers(0, len(mats), n_tri).astype(np.uint32)
    return S.SceneArrays(verts=v, indices=np.asarray(faces, np.uint32), mat_ids=mat_ids, materials=mats, name="soup")


@pytest.mark.parametrize("seed", range(8))
def test_random_soups_and_cameras(api, port, seed):
    rng = np.random.default_rng(100 + seed)
    n_tri = int(rng.choice([1, 2, 7, 31, 48, 64, 65, 200]))          # <= 64: candidate lists on; above: traversal only
    arr = _random_soup(rng, n_tri)
    gs, osc = api.Scene(arr), O.OracleScene(port, arr)
    try:
        for view in range(3):
            eye = rng.uniform([-300, -100, -900], [850, 700, 800])
            lookat = rng.uniform([100, 100, 100], [450, 450, 450])
            up = rng.normal(size=3)
            fovy = float(rng.uniform(10, 110))
            w, h = int(rng.integers(9, 120)), int(rng.integers(9, 80))
            pp = _with_camera(api, api.make_params(w, h, 3, 4, frame=view, scene=gs), eye, lookat, up, fovy)
            g, o = gs.debug_primary(pp), osc.render(pp, math_mode=1, use_bvh=True, probes=True)
            assert np.array_equal(g["prim"], o["prim"]), (seed, view)
            assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32)), (seed, view)
            p = _with_camera(api, api.make_params(w, h, 5, 9, frame=view, nee=True, importance=bool(view & 1), scene=gs),
                             eye, lookat, up, fovy)
            gr, orr = gs.launch_host(p), osc.render(p, math_mode=1, use_bvh=True)
            # (frames > 0 fold into a zero previous accumulation on both sides)
            assert np.array_equal(gr["accum"].view(np.uint32), orr["accum"].view(np.uint32)), (seed, view)
            assert np.array_equal(gr["fb"], orr["fb"]), (seed, view)
    finally:
        gs.destroy()


# ---- node fetch variants (pt_scene_set_traversal): same pixels ------------------------------------------------------------
TRAVERSALS = ((2, "binary"), (3, "binary-quant"), (4, "wide4-quant"), (1, "binary"))   # global exact, quantised, 4-wide, auto placement


@pytest.mark.parametrize("name,w,h,spp,md", [("cornell", 160, 90, 6, 16), ("material", 200, 112, 6, 28), ("tess40", 160, 90, 4, 8),
                                             ("tess300", 128, 72, 3, 8)])
def test_traversal_variants_same_pixels(api, port, scenes_cache, name, w, h, spp, md):
    """pt_scene_set_traversal: exact nodes read from global memory (also for scenes that would be staged in shared
    memory), quantised 32-byte nodes, the default placement and the oracle give the same primary hits and the same
    pixels, bit for bit."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    p = api.make_params(w, h, spp, md, frame=1, nee=True, importance=True, scene=gs)
    prev = np.random.default_rng(3).random((h, w, 4), dtype=np.float32)
    o = osc.render(p, accum=prev.copy(), math_mode=1, use_bvh=True, probes=True)
    try:
        for trav, expect in TRAVERSALS:
            gs.set_traversal(trav)
            g = gs.launch_host(p, accum=prev.copy())
            assert gs.last_launch_info()["traversal"] == expect
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), trav
            assert np.array_equal(g["fb"], o["fb"]), trav
            d = gs.debug_primary(p)
            assert np.array_equal(d["prim"], o["prim"]) and np.array_equal(d["t"].view(np.uint32), o["t"].view(np.uint32))
    finally:
        gs.set_traversal(0)


def test_traversal_variants_single_triangle_and_soups(api, port):
    rng = np.random.default_rng(77)
    for n_tri in (1, 2, 3, 5, 70, 333):
        arr = _random_soup(rng, n_tri)
        gs, osc = api.Scene(arr), O.OracleScene(port, arr)
        try:
            p = api.make_params(64, 40, 4, 6, frame=0, nee=True, scene=gs)
            o = osc.render(p, math_mode=1, use_bvh=True)
            for trav, expect in TRAVERSALS:
                if n_tri < 2 and trav in (3, 4):
                    expect = "binary"                         # a single triangle has no quantised node array: exact nodes
                gs.set_traversal(trav)
                g = gs.launch_host(p)
                assert gs.last_launch_info()["traversal"] == expect
                assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), (n_tri, trav)
            if n_tri >= 2:                                   # the three trees of a tiny scene: same pixels, PLOC == oracle's
                gp, op = gs.bvh_export_ploc(), osc.bvh_ploc()
                a, b = gp["nodes"].copy(), op["nodes"].copy()
                a[:, [11, 15]] = 0
                b[:, [11, 15]] = 0
                assert gp["depth"] == op["depth"] and np.array_equal(a, b), n_tri
                for trav in (0, 3, 4):                       # exact, quantised and 4-wide nodes of each tree
                    gs.set_traversal(trav)
                    for tree in (1, 2, 3):
                        gs.set_tree(tree)
                        g = gs.launch_host(p)
                        assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), (n_tri, tree, trav)
                gs.set_tree(0)
                gs.set_traversal(0)
            # counted launches from global memory report the specified traversal's work counters
            gs.set_traversal(2)
            gs.enable_counters(True)
            gs.counters(reset=True)
            g = gs.launch_host(p)
            assert gs.counters(reset=True) == o["counters"]
            gs.enable_counters(False)
        finally:
            gs.destroy()


# ---- sweep-SAH tree (pt_scene_set_tree): same leaves, better topology, same pixels ------------------------------------
@pytest.mark.parametrize("name", ["cornell", "material", "tess40"])
def test_sah_tree_bit_exact_and_selectable(api, port, scenes_cache, name):
    arr, gs, osc = _get(scenes_cache, api, port, name)
    g, o = gs.bvh_export_sah(), osc.bvh_sah()
    gn, on = g["nodes"].copy(), o["nodes"].copy()
    gn[:, [11, 15]] = 0
    on[:, [11, 15]] = 0
    assert g["depth"] == o["depth"] and np.array_equal(gn, on)
    assert gs.tree_in_use() == "sah" == osc.set_tree(0)                     # auto: 2..8192 triangles
    org, d = _rays(arr, 100_000, 17)
    w, h = 160, 90
    p = api.make_params(w, h, 4, 12, frame=2, nee=True, importance=True, scene=gs)
    prev = np.random.default_rng(4).random((h, w, 4), dtype=np.float32)
    results = {}
    try:
        for tree in (2, 3, 1):
            gs.set_tree(tree)
            name_t = osc.set_tree(tree)
            assert gs.tree_in_use() == name_t
            gp, gt, ga, gc = gs.trace(org, d, 0.01, 1e16, want_any=True)
            op, ot, oa, oc = osc.trace(org, d, 0.01, 1e16, use_bvh=True, want_any=True)
            assert np.array_equal(gp, op) and np.array_equal(gt.view(np.uint32), ot.view(np.uint32)) and np.array_equal(ga, oa)
            assert gc.tolist() == oc.tolist()                                   # work counters of THAT tree
            gs.enable_counters(True)
            gs.counters(reset=True)
            gr = gs.launch_host(p, accum=prev.copy())
            orr = osc.render(p, accum=prev.copy(), math_mode=1, use_bvh=True)
            assert gs.counters(reset=True) == orr["counters"]
            gs.enable_counters(False)
            assert np.array_equal(gr["accum"].view(np.uint32), orr["accum"].view(np.uint32))
            un = gs.launch_host(p, accum=prev.copy())                           # the uncounted (timed) kernel
            assert np.array_equal(un["accum"].view(np.uint32), orr["accum"].view(np.uint32)) and np.array_equal(un["fb"], orr["fb"])
            results[tree] = (gr["accum"], gc.tolist())
        assert np.array_equal(results[1][0].view(np.uint32), results[2][0].view(np.uint32))
        assert np.array_equal(results[1][0].view(np.uint32), results[3][0].view(np.uint32))
        assert results[2][1][0] < results[1][1][0] and results[3][1][0] < results[1][1][0]   # fewer node visits than the LBVH
    finally:
        gs.enable_counters(False)
        gs.set_tree(0)
        osc.set_tree(0)


def test_sah_tree_limits(api, port):
    """No SAH tree above 8192 triangles (the PLOC tree is in use there; asking for SAH is an error)."""
    big = S.tessellated_scene(70)
    gs = api.Scene(big)
    try:
        assert gs.tree_in_use() == "ploc"
        with pytest.raises(api.PtError):
            gs.set_tree(2)
    finally:
        gs.destroy()


@pytest.mark.parametrize("name", ["cornell", "material", "tess40", "tess300"])
def test_ploc_tree_bit_exact(api, port, scenes_cache, name):
    """The GPU-built PLOC tree (lbvh.cu ploc_* kernels) equals the oracle's serial restatement, word for word."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    g, o = gs.bvh_export_ploc(), osc.bvh_ploc()
    gn, on = g["nodes"].copy(), o["nodes"].copy()
    gn[:, [11, 15]] = 0
    on[:, [11, 15]] = 0
    assert g["depth"] == o["depth"]
    assert np.array_equal(gn, on)
    assert gs.tree_in_use() == osc.set_tree(0) == ("sah" if arr.n_tri <= 8192 else "ploc")


@pytest.mark.parametrize("name,w,h,spp,md", [("material", 160, 90, 6, 28), ("tess40", 160, 90, 4, 8), ("tess300", 128, 72, 3, 8)])
@pytest.mark.parametrize("levels", [2, 5, 0])
def test_stack_spill_same_pixels(api, port, scenes_cache, name, w, h, spp, md, levels):
    """Scenes read from global memory keep only the first levels of a path's traversal stack in shared memory (12 by
    default) and spill deeper ones to global memory; forced down to 2 and 5 levels (and with no spill at all) the pixels
    and the work counters are still the oracle's."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    p = api.make_params(w, h, spp, md, frame=1, nee=True, importance=True, scene=gs)
    o = osc.render(p, math_mode=1, use_bvh=True)
    gs.set_tune(f"stack_smem={levels}")
    try:
        gs.set_traversal(2)                                       # material (672 triangles) would be staged in shared memory
        g = gs.launch_host(p)
        assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32))
        gs.enable_counters(True)
        gs.counters(reset=True)
        g = gs.launch_host(p)
        assert gs.counters(reset=True) == o["counters"]
        assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32))
    finally:
        gs.enable_counters(False)
        gs.set_traversal(0)
        gs.set_tune("")


def _degenerate_scene(kind):
    """Inputs that stress tie-breaking in the tree builders: equal Morton keys, equal SAH costs, equal PLOC areas."""
    rng = np.random.default_rng(3)
    tri = np.array([[100, 100, 300], [450, 120, 320], [280, 480, 340]], np.float32)
    if kind == "copies":                      # 300 copies of one triangle
        v = np.tile(tri, (300, 1))
    elif kind == "stack":                     # 200 parallel copies, shifted along z by multiples of an ulp-sized step
        v = np.concatenate([tri + np.array([0, 0, 0.5 * k], np.float32) for k in range(200)])
    else:                                     # zero-area triangles (points, segments) mixed with real ones
        v = rng.uniform(50, 500, (450, 3)).astype(np.float32)
        v[0:90] = v[0]                        # 30 triangles collapsed to one point
        v[91::3][:40] = v[90::3][:40]         # segments
    v4 = np.ones((v.shape[0], 4), np.float32)
    v4[:, :3] = v
    n = v.shape[0] // 3
    mats = [dict(name="white", Kd=(0.8, 0.8, 0.8), Ke=(0, 0, 0)), dict(name="RefractiveThing", Kd=(1, 1, 1), Ke=(0, 0, 0), Ni=1.5)]
    return S.SceneArrays(verts=v4, indices=np.arange(3 * n, dtype=np.uint32).reshape(n, 3),
                         mat_ids=(np.arange(n) % 2).astype(np.uint32), materials=mats, name=kind)


@pytest.mark.parametrize("kind", ["copies", "stack", "flat"])
def test_degenerate_scenes_all_trees(api, port, kind):
    arr = _degenerate_scene(kind)
    gs, osc = api.Scene(arr), O.OracleScene(port, arr)
    try:
        for exp_g, exp_o in ((gs.bvh_export_sah(), osc.bvh_sah()), (gs.bvh_export_ploc(), osc.bvh_ploc())):
            a, b = exp_g["nodes"].copy(), exp_o["nodes"].copy()
            a[:, [11, 15]] = 0
            b[:, [11, 15]] = 0
            assert exp_g["depth"] == exp_o["depth"] and np.array_equal(a, b)
        p = api.make_params(96, 54, 3, 12, frame=0, nee=True, scene=gs)
        for tree in (0, 1, 2, 3):
            for trav in (0, 2, 3, 4):
                used = osc.set_tree(tree)
                try:
                    gs.set_tree(tree)
                except api.PtError:                  # a tree deeper than 64 levels (300 copies of one triangle) is refused
                    assert used == "lbvh" and tree in (2, 3)
                    gs.set_tree(1)
                gs.set_traversal(trav)
                assert gs.tree_in_use() == used
                g, o = gs.launch_host(p), osc.render(p, math_mode=1, use_bvh=True)
                assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), (kind, tree, trav)
    finally:
        gs.destroy()


def _collapse_wide(q, step):
    """The 4-wide collapse restated (lbvh.cu build_wide_nodes): breadth first; per node the two children of a binary node,
    then twice the internal child with the largest surface area (first one on ties) replaced by its own two children."""
    sx, sy, sz = (float(v) for v in step)

    def child(node, k):
        w0, w1, w2 = (int(v) for v in q[node, 3 * k:3 * k + 3])
        lx, ly, lz, hx, hy, hz = w0 & 0xFFFF, w0 >> 16, w1 & 0xFFFF, w1 >> 16, w2 & 0xFFFF, w2 >> 16
        dx, dy, dz = sx * (float(hx) - float(lx)), sy * (float(hy) - float(ly)), sz * (float(hz) - float(lz))
        ref = int(np.int32(q[node, 6 + k]))
        return [lx | (hx << 16), ly | (hy << 16), lz | (hz << 16)], ref, dx * dy + dy * dz + dz * dx

    binary_of, level, out = [0], [1], []
    wi = 0
    while wi < len(binary_of):
        ch = [child(binary_of[wi], 0), child(binary_of[wi], 1)]
        while len(ch) < 4:
            best = -1
            for k, c in enumerate(ch):
                if c[1] >= 0 and (best < 0 or c[2] > ch[best][2]):
                    best = k
            if best < 0:
                break
            b = ch[best][1]
            ch[best] = child(b, 0)
            ch.append(child(b, 1))
        rec = [0] * 16
        for k in range(4):
            if k >= len(ch):
                rec[3 * k:3 * k + 3] = [0xFFFF] * 3
                continue
            rec[3 * k:3 * k + 3] = ch[k][0]
            if ch[k][1] >= 0:
                rec[12 + k] = len(binary_of)
                binary_of.append(ch[k][1])
                level.append(level[wi] + 1)
            else:
                rec[12 + k] = ch[k][1] & 0xFFFFFFFF
        for k in range(len(ch), 4):
            rec[12 + k] = rec[12]
        out.append(rec)
        wi += 1
    return np.array(out, dtype=np.uint32), max(level)


@pytest.mark.parametrize("name", ["cornell", "material", "tess40"])
def test_wide_nodes_are_the_collapse_of_the_quantised_tree(api, port, scenes_cache, name):
    """64-byte 4-wide nodes (pt_scene_set_traversal 4): the GPU build equals the restated collapse record for record, for every
    tree; every leaf is referenced exactly once and every wide node but the root exactly once."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    try:
        for tree in (1, 2, 3):
            gs.set_tree(tree)
            qd = gs.bvh_export_quant()
            wd = gs.bvh_export_wide()
            want, depth = _collapse_wide(qd["q"], qd["step"])
            assert wd["w"].shape == want.shape and wd["depth"] == depth, (tree, wd["w"].shape, want.shape, wd["depth"], depth)
            assert np.array_equal(wd["w"], want), tree
            refs, boxes = wd["w"][:, 12:].astype(np.int64), wd["w"][:, :12].reshape(-1, 4, 3)
            real = ~np.all(boxes == 0xFFFF, axis=2)                       # not an empty slot
            signed = np.where(refs >= 2**31, refs - 2**32, refs)
            leaves = np.sort(~signed[real & (signed < 0)])
            assert np.array_equal(leaves, np.arange(arr.n_tri)), tree
            inner = np.sort(signed[real & (signed >= 0)])
            assert np.array_equal(inner, np.arange(1, len(want))), tree
    finally:
        gs.set_tree(0)


@pytest.mark.parametrize("name", ["cornell", "material", "tess40", "tess300"])
def test_quantised_nodes_contain_the_exact_boxes(api, port, scenes_cache, name):
    """32-byte nodes (pt_scene_set_traversal 3): every dequantised plane -- fma(2^23 + q, step, base), the traversal's expression,
    evaluated here in float64 and rounded once like a float32 fma -- is on the outside of the exact plane, the child
    references are the exact node's, and a plane is never more than one grid step away."""
    arr, gs, osc = _get(scenes_cache, api, port, name)
    try:
        for tree, export in ((1, gs.bvh_export), (2, gs.bvh_export_sah), (3, gs.bvh_export_ploc)):
            if tree == 2 and arr.n_tri > 8192:
                continue
            gs.set_tree(tree)
            exact = export()["nodes"]
            qd = gs.bvh_export_quant()
            q, lo, step = qd["q"], qd["lo"].astype(np.float64), qd["step"].astype(np.float64)
            f = exact.view(np.float32)
            assert np.array_equal(q[:, 6], exact[:, 3]) and np.array_equal(q[:, 7], exact[:, 7])
            w = q[:, :6]
            planes = np.stack([w[:, 0] & 0xFFFF, w[:, 0] >> 16, w[:, 1] & 0xFFFF, w[:, 1] >> 16, w[:, 2] & 0xFFFF, w[:, 2] >> 16,
                               w[:, 3] & 0xFFFF, w[:, 3] >> 16, w[:, 4] & 0xFFFF, w[:, 4] >> 16, w[:, 5] & 0xFFFF, w[:, 5] >> 16], axis=1)
            axis = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2, 0, 1, 2])
            # (2^23 + q) * step is exact in float64 (24 x 24 bits), the sum is rounded once to float32: the value of the fused
            # multiply-add; `lo` is the addend of that expression (grid origin - 2^23 * step)
            deq = ((planes.astype(np.float64) + 8388608.0) * step[axis] + lo[axis]).astype(np.float32)
            ex = np.stack([f[:, 0], f[:, 1], f[:, 2], f[:, 4], f[:, 5], f[:, 6], f[:, 8], f[:, 9], f[:, 10], f[:, 12], f[:, 13], f[:, 14]], axis=1)
            is_lo = np.array([1, 1, 1, 0, 0, 0, 1, 1, 1, 0, 0, 0], bool)
            assert np.all(deq[:, is_lo] <= ex[:, is_lo]) and np.all(deq[:, ~is_lo] >= ex[:, ~is_lo]), tree
            assert np.all(np.abs(deq.astype(np.float64) - ex) <= 1.001 * step[axis] + 1e-4), tree
    finally:
        gs.set_tree(0)


# ---- sample teams (wavefront.cu "Sample teams"): several slots work on one pixel, ordered commit ------------------------------
TEAM_CASES = [
    # scene, w, h, spp, maxDepth, tune -- few pixels and many samples: almost every sample is rendered by a team
    ("cornell", 40, 24, 64, 16, "team=2"),             # (team=2: also for scenes staged in shared memory, where auto leaves it off)
    ("cornell", 9, 5, 200, 16, "team=2"),              # fewer pixels than columns
    ("material", 48, 27, 48, 28, "team=2"),            # long refractive / conductor paths: the commit frontier lags
    ("material", 48, 27, 48, 28, "team=2,team_min=40"),  # pixels switch late or not at all
    ("tess40", 32, 18, 32, 8, ""),                     # scene in global memory: teams by default
    ("cornell", 24, 16, 96, 200, "team=2"),            # deep logs: 203 entries per sample, a window of 16 samples (parking)
    ("cornell", 40, 24, 64, 16, "team=2,paths32=4,block=128"),
    ("cornell", 40, 24, 64, 16, "team=0"),             # the sequential chain, for reference
    ("material", 64, 36, 24, 28, "team=2,msort=1"),    # SHADE batches sorted by BSDF type (three queues)
    ("material", 64, 36, 24, 28, "team=0,msort=1"),
]


@pytest.mark.parametrize("case", TEAM_CASES, ids=lambda c: "-".join(str(x) for x in c))
def test_sample_teams_bit_exact(api, port, scenes_cache, case):
    name, w, h, spp, md, tune = case
    arr, gs, osc = _get(scenes_cache, api, port, name)
    gs.set_tune(tune)
    try:
        prev = None
        fresh = lambda: None if prev is None else prev.copy()       # both sides fold in place: give each its own copy
        for f in range(2):
            p = api.make_params(w, h, spp, md, frame=f, nee=True, importance=True, scene=gs)
            o = osc.render(p, accum=fresh(), math_mode=1, use_bvh=True)
            g = gs.launch_host(p, accum=fresh())
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), f"frame {f}"
            assert np.array_equal(g["fb"], o["fb"]), f"frame {f}"
            gs.enable_counters(True)                 # the counted kernel forms its own teams: same work, same pixels
            gs.counters(reset=True)
            g = gs.launch_host(p, accum=fresh())
            assert gs.counters(reset=True) == o["counters"]
            gs.enable_counters(False)
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), f"frame {f} (counted)"
            prev = o["accum"]
        gs.check_error()
    finally:
        gs.enable_counters(False)
        gs.set_tune("")


def test_small_share_of_a_frame_uses_every_sm(api, port, scenes_cache):
    """One 1/8 tile share of a 16:9 frame at 128 spp: with teams the launch spreads over the whole GPU (64 pixels per CTA are
    enough) instead of one CTA per 1024 pixels, and the pixels are the oracle's."""
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    w, h = 256, 144
    acc, fb = _dev_buffers(w, h)
    p = api.make_params(w, h, 128, 16, frame=0, nee=True, importance=True, scene=gs)
    gs.set_tune("team=2")
    _launch_dev(api, gs, p, acc, fb, rank=3, n=8)
    grid_team = gs.last_launch_info()["grid"]
    got = acc.cpu().numpy().copy()
    gs.set_tune("team=0")
    try:
        acc.zero_()
        _launch_dev(api, gs, p, acc, fb, rank=3, n=8)
        grid_solo = gs.last_launch_info()["grid"]
        assert np.array_equal(acc.cpu().numpy().view(np.uint32), got.view(np.uint32))
    finally:
        gs.set_tune("")
    assert grid_solo == 5 and grid_team == 72, (grid_solo, grid_team)     # 4608 owned pixels: / 1024 slots vs / 64 pixels
    from acgpathtracing_b200.multirank import owned_mask
    mine = owned_mask(w, h, 3, 8)
    o = osc.render(p, math_mode=1, use_bvh=True)
    assert np.array_equal(got[mine].view(np.uint32), o["accum"][mine].view(np.uint32))


# ---- the lock-free scheduler under perturbation (stands in for compute-sanitizer racecheck, which this GPU pool does not
# allow): the image and the work counters must not depend on how paths are batched -------------------------------------------
def test_schedule_perturbation(api, port, scenes_cache):
    arr, gs, osc = _get(scenes_cache, api, port, "material")
    w, h, spp, md = 160, 90, 24, 28
    p = api.make_params(w, h, spp, md, frame=1, nee=True, importance=True, scene=gs)
    prev = np.random.default_rng(5).random((h, w, 4), dtype=np.float32)
    o = osc.render(p, accum=prev.copy(), math_mode=1, use_bvh=True)
    rng = np.random.default_rng(2024)
    tried = []
    try:
        for trial in range(20):
            block = int(rng.choice([0, 128, 256, 384, 512, 768]))
            paths32 = int(rng.choice([0, 4, 8, 16, 24]))
            if paths32 and block > 32 * paths32:
                block = 32 * paths32
            spec = dict(block=block, paths32=paths32, node_votes=int(rng.integers(1, 9)), node_keep8=int(rng.integers(0, 9)),
                        full_cols=int(rng.integers(1, 33)), sticky=int(rng.integers(0, 9)), tail_cols=int(rng.choice([0, 8, 24, 64, 160])),
                        idle_ns=int(rng.choice([0, 16, 64, 512])), idle_shift=int(rng.integers(0, 6)), stack_smem=int(rng.choice([0, 2, 5, 12])),
                        team=int(rng.choice([0, 1, 2, 2])), team_min=int(rng.choice([1, 2, 8, 20])), msort=int(rng.choice([0, 1])))
            tune = ",".join(f"{k}={v}" for k, v in spec.items())
            tried.append(tune)
            gs.set_tune(tune)
            gs.set_traversal(int(rng.choice([1, 2, 3, 4])))       # staged, global exact, global quantised, global 4-wide
            gs.set_pipeline(2)
            g = gs.launch_host(p, accum=prev.copy())
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), tune
            assert np.array_equal(g["fb"], o["fb"]), tune
            gs.enable_counters(True)
            gs.counters(reset=True)
            g = gs.launch_host(p, accum=prev.copy())
            assert gs.counters(reset=True) == o["counters"], tune
            gs.enable_counters(False)
            assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32)), tune
        gs.check_error()
    finally:
        gs.enable_counters(False)
        gs.set_tune("")
        gs.set_traversal(0)
        gs.set_pipeline(0)
    assert len(set(tried)) >= 16


def test_soak_200_launches_no_watchdog(api, port, scenes_cache):
    """200 launches of a team-heavy configuration in a row: every one equal to the first, the watchdog flag never set."""
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    w, h = 64, 36
    acc, fb = _dev_buffers(w, h)
    p = api.make_params(w, h, 48, 16, frame=0, nee=True, importance=True, scene=gs)
    gs.set_tune("team=2")
    _launch_dev(api, gs, p, acc, fb)
    first = acc.clone()
    o = osc.render(p, math_mode=1, use_bvh=True)
    assert np.array_equal(first.cpu().numpy().view(np.uint32), o["accum"].view(np.uint32))
    import torch
    for i in range(200):
        acc.zero_()
        _launch_dev(api, gs, p, acc, fb)
        assert torch.equal(acc.view(torch.int32), first.view(torch.int32)), i
    gs.check_error()
    gs.set_tune("")


# ---- launches the wavefront record cannot hold fall back to pipeline 1 ----------------------------------------
def test_max_depth_above_255_terminates_and_matches_oracle(api, port):
    """maxDepth = 300 on a closed refractive box with Kd = 1: total internal reflection keeps Russian roulette from ever
    ending a path, so only `depth >= maxDepth` (pathTracerPrograms.cu:767) does.  The wavefront record packs the depth
    into 8 bits; such launches run on pipeline 1 and equal the oracle."""
    c = np.array([[100, 100, 200], [400, 100, 200], [400, 400, 200], [100, 400, 200],
                  [100, 100, 500], [400, 100, 500], [400, 400, 500], [100, 400, 500]], np.float32)
    quads = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 3, 7, 4)]
    idx = np.array([t for q in quads for t in ((q[0], q[1], q[2]), (q[0], q[2], q[3]))], np.uint32)
    v4 = np.ones((8, 4), np.float32)
    v4[:, :3] = c
    mats = [dict(name="RefractiveBox", Kd=(1, 1, 1), Ke=(0, 0, 0), Ni=1.5)]
    arr = S.SceneArrays(verts=v4, indices=idx, mat_ids=np.zeros(12, np.uint32), materials=mats, name="glassbox")
    gs, osc = api.Scene(arr, background=(0.2, 0.3, 0.4)), O.OracleScene(port, arr, background=(0.2, 0.3, 0.4))
    try:
        p = api.make_params(48, 32, 2, 300, frame=0, nee=True, importance=True, scene=gs)
        g = gs.launch_host(p)
        o = osc.render(p, math_mode=1, use_bvh=True)
        assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32))
        assert gs.last_launch_info()["block"] == 256                       # render_persistent_kernel
        p255 = api.make_params(48, 32, 2, 255, frame=0, nee=True, importance=True, scene=gs)
        g = gs.launch_host(p255)
        o = osc.render(p255, math_mode=1, use_bvh=True)
        assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32))
        assert gs.last_launch_info()["block"] != 256                       # the wavefront kernel holds 255
        gs.set_pipeline(2)
        with pytest.raises(api.PtError):                                   # asked for the wavefront explicitly: refused, not wrong
            gs.launch_host(p)
    finally:
        gs.destroy()


def test_image_wider_than_65535_pixels(api, port, scenes_cache):
    """The wavefront record packs a pixel into 16 + 16 bits; a 70000 x 2 image runs on pipeline 1 and equals the oracle."""
    arr, gs, osc = _get(scenes_cache, api, port, "cornell")
    p = api.make_params(70000, 2, 1, 4, frame=0, nee=True, importance=True, scene=gs)
    g = gs.launch_host(p)
    o = osc.render(p, math_mode=1, use_bvh=True)
    assert np.array_equal(g["accum"].view(np.uint32), o["accum"].view(np.uint32))
    assert gs.last_launch_info()["block"] == 256


# ---- candidate lists: a million primary rays against brute force ---------------------------------------------------------------
def _adversarial_soup(rng, eye, lookat, n_tri):
    """Triangles placed to stress the side-plane test of pixel_candidates (wavefront.cu): tiny ones (smaller than a pixel's
    footprint), slivers along the view direction, triangles with an edge or a vertex on the view axis, and large ones that
    leave the frustum -- at distances from 1 to 4000 units from the eye."""
    axis = (lookat - eye) / np.linalg.norm(lookat - eye)
    side = np.cross(axis, rng.normal(size=3))
    side /= np.linalg.norm(side)
    up = np.cross(side, axis)
    verts = []
    for k in range(n_tri):
        dist = float(np.exp(rng.uniform(np.log(1.0), np.log(4000.0))))
        centre = eye + axis * dist + (side * rng.normal() + up * rng.normal()) * dist * 0.15
        size = dist * float(np.exp(rng.uniform(np.log(1e-4), np.log(0.5))))
        kind = k % 4
        if kind == 0:       # isotropic, any size from sub-pixel to frustum-filling
            e1, e2 = rng.normal(size=3) * size, rng.normal(size=3) * size
        elif kind == 1:     # sliver pointing at the eye
            e1 = axis * size * 5 + rng.normal(size=3) * size * 0.01
            e2 = rng.normal(size=3) * size * 0.02
        elif kind == 2:     # a vertex exactly on the view axis
            centre = eye + axis * dist
            e1, e2 = (side * rng.normal() + up * rng.normal()) * size, (side * rng.normal() + up * rng.normal()) * size
        else:               # edge-on: nearly parallel to the rays through it
            e1 = (centre - eye) / dist * size + rng.normal(size=3) * size * 1e-3
            e2 = rng.normal(size=3) * size
        verts += [centre, centre + e1, centre + e2]
    v = np.ones((len(verts), 4), np.float32)
    v[:, :3] = np.asarray(verts, np.float32)
    mats = [dict(name="white", Kd=(0.8, 0.8, 0.8), Ke=(0, 0, 0))]
    return S.SceneArrays(verts=v, indices=np.arange(3 * n_tri, dtype=np.uint32).reshape(n_tri, 3), mat_ids=np.zeros(n_tri, np.uint32),
                         materials=mats, name="adversarial")


def test_candidate_lists_never_drop_a_hit(api, port):
    """pixel_candidates keeps a triangle unless all three vertices are outside one side plane of the pixel's ray pyramid by
    more than 2e-4 of their distance.  In exact arithmetic a camera ray of the pixel cannot hit such a triangle with margin 0;
    float error in the plane normals, in the ray direction and in the triangle test itself is bounded by a few 1e-6 of that
    distance ("margin"), two orders below the margin.  Checked here the blunt way: more than a million primary
    rays through scenes built to sit on those planes, against brute force over every triangle (the oracle without a tree)."""
    rng = np.random.default_rng(4242)
    rays = 0
    for trial in range(36):
        eye = rng.uniform([-500, -500, -1500], [1000, 1000, 500])
        lookat = eye + rng.normal(size=3) * 300
        n_tri = int(rng.choice([3, 16, 40, 64]))
        arr = _adversarial_soup(rng, eye, lookat, n_tri)
        gs, osc = api.Scene(arr), O.OracleScene(port, arr)
        try:
            w, h = [(320, 200), (64, 64), (1024, 40), (9, 7)][trial % 4]
            fovy = float(rng.uniform(2.0, 120.0))                        # from a telephoto sliver of the scene to a fisheye
            for frame in range(3):
                pp = _with_camera(api, api.make_params(w, h, 1, 1, frame=frame, scene=gs), eye, lookat, rng.normal(size=3), fovy)
                g = gs.debug_primary(pp)
                o = osc.render(pp, math_mode=1, use_bvh=False, probes=True, want_fb=False)     # brute force
                assert np.array_equal(g["prim"], o["prim"]), (trial, frame)
                assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32)), (trial, frame)
                rays += w * h
            assert arr.n_tri <= 64                                       # the size up to which the kernel builds candidate lists
        finally:
            gs.destroy()
    assert rays >= 1_000_000
