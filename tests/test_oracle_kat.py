"""Known-answer vectors captured from the reference's own functions against the
oracle port -- and against the reference itself (oracle/_ref) when it is built."""
import ctypes as C

import numpy as np
import pytest

F = np.float32


def _close(a, b, rel=2e-7, abs_=1e-12):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.all(np.abs(a - b) <= rel * np.abs(b) + abs_)


def test_tea_and_lcg(port):
    # cuda/random.h:32-67
    assert port.ptor_tea4(0, 0) == 1576399551
    assert port.ptor_tea4(262143, 3) == 2391278879
    s = C.c_uint32(1576399551)
    got = [(port.ptor_rnd(C.byref(s)), s.value) for _ in range(3)]
    want = [(0.294449925, 3041616146), (0.695515215, 4105309513), (0.897309542, 3873814036)]
    for (g, gs), (w, ws) in zip(got, want):
        assert gs == ws
        assert F(g) == F(w)


def test_rnd_is_exact_24bit(port):
    s = C.c_uint32(7)
    for _ in range(1000):
        r = port.ptor_rnd(C.byref(s))
        assert r == (s.value & 0xFFFFFF) / 16777216.0


def test_camera_uvw(port):
    # sutil/Camera.cpp:33-45 with PathTracerMain.cpp:228-233
    import oracle_api as O
    for (w, h), u in (((512, 512), -387.817566), ((1920, 1080), -689.45343), ((3840, 2160), -689.45343)):
        p = O.make_params(port, w, h, 1, 1)
        assert _close(p.cameraU.tolist(), [u, 0, 0], rel=1e-7)
        assert _close(p.cameraV.tolist(), [0, 387.817566, 0], rel=1e-7)
        assert p.cameraW.tolist() == [0.0, 0.0, 1230.0]


@pytest.mark.parametrize("mode", [0, 1])
def test_samplers(port, mode):
    out = np.zeros(3, F)
    tol = 2e-7 if mode == 0 else 6e-7    # det mode: within 2 ulp of glibc
    port.ptor_cosine_sample_hemisphere(mode, 0.25, 0.5, out)     # pathTracerPrograms.cu:341-353
    assert _close(out, [-0.866025448, -7.57103464e-08, 0.49999997], rel=tol, abs_=2e-7)
    port.ptor_uniform_sample_hemisphere(mode, 0.25, 0.5, out)    # :368-380
    assert _close(out, [-0.968245864, -8.46467429e-08, 0.25], rel=tol, abs_=2e-7)
    port.ptor_sample_ggx(mode, 0.25, 0.5, 0.2, np.array([0, 1, 0], F), out)   # :455-476
    assert _close(out, [8.57250537e-09, 0.980580688, 0.19611606], rel=tol, abs_=2e-7)
    n = np.array([0.03, 0, 0.99955], np.float64)
    n = (n / np.linalg.norm(n)).astype(F)
    port.ptor_sample_ggx(mode, 0.25, 0.5, 0.2, n, out)           # float abs() branch
    assert _close(out, [0.225445196, 8.57250448e-09, 0.97425586], rel=tol, abs_=2e-7)


def test_fresnel(port):
    assert F(port.ptor_fr_dielectric(1.0, 1.0, 1.5)) == F(0.0400000028)      # :534-559
    assert F(port.ptor_fr_dielectric(0.5, 1.0, 1.5)) == F(0.0891867131)
    assert port.ptor_fr_dielectric(-0.5, 1.0, 1.5) == 1.0                    # total internal reflection
    out = np.zeros(3, F)
    eta = np.array([1.45, 0.7, 1.55], F)
    k = np.array([3.0, 2.2, 3.5], F)
    port.ptor_fresnel_conductor(1.0, eta, k, out)                            # :494-510
    assert _close(out, [0.538148642, 0.604049921, 0.591056824])
    port.ptor_fresnel_conductor(0.5, eta, k, out)
    assert _close(out, [0.699393451, 0.71808517, 0.744461894])


def test_onb(port):
    t, b = np.zeros(3, F), np.zeros(3, F)
    port.ptor_onb(np.array([0, 1, 0], F), t, b)                              # :54-85
    assert t.tolist() == [-1.0, 0.0, 0.0]
    assert b.tolist() == [0.0, 0.0, 1.0] and np.signbit(b[1])


@pytest.mark.parametrize("mode", [0, 1])
def test_all_miss_image(port, mode):
    """__raygen__rg on a 2x2 all-miss image, background (.25,.5,.75), spp 2."""
    import oracle_api as O
    from acgpathtracing_b200.scenes import SceneArrays
    empty = SceneArrays(np.ones((3, 4), F), np.zeros((0, 3), np.uint32), np.zeros(0, np.uint32),
                        [dict(name="m", Kd=(1, 1, 1), Ke=(0, 0, 0))])
    sc = O.OracleScene(port, empty, background=(0.25, 0.5, 0.75))
    p = O.make_params(port, 2, 2, 2, 3)
    r = sc.render(p, math_mode=mode, use_bvh=True)
    assert np.array_equal(r["accum"], np.broadcast_to(np.array([0.25, 0.5, 0.75, 1.0], F), (2, 2, 4)))
    assert np.array_equal(r["fb"], np.broadcast_to(np.array([137, 188, 225, 255], np.uint8), (2, 2, 4)))


def test_struct_sizes(port):
    assert port.ptor_sizeof_params() == 168
    assert port.ptor_sizeof_material() == 40


# ---- the same vectors straight out of the reference's code (oracle/_ref) -------------------------------
def test_ref_kat(ref, port):
    assert ref.ptref_tea4(0, 0) == 1576399551 and ref.ptref_tea4(262143, 3) == 2391278879
    a, b = np.zeros(3, F), np.zeros(3, F)
    rng = np.random.default_rng(0)
    for _ in range(2000):
        u1, u2 = F(rng.random()), F(rng.random())
        ref.ptref_cosine_sample_hemisphere(u1, u2, a)
        port.ptor_cosine_sample_hemisphere(0, u1, u2, b)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        ref.ptref_uniform_sample_hemisphere(u1, u2, a)
        port.ptor_uniform_sample_hemisphere(0, u1, u2, b)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        n = rng.normal(size=3)
        n = (n / np.linalg.norm(n)).astype(F)
        ref.ptref_sample_ggx(u1, u2, 0.2, n, a)
        port.ptor_sample_ggx(0, u1, u2, 0.2, n, b)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        c = F(rng.uniform(-1, 1))
        assert F(ref.ptref_fr_dielectric(c, 1.0, 1.5)) == F(port.ptor_fr_dielectric(c, 1.0, 1.5))
        i = rng.normal(size=3)
        i = (i / np.linalg.norm(i)).astype(F)
        ra = ref.ptref_refract(i, n, 1.5, a)
        rb = port.ptor_refract(i, n, 1.5, b)
        assert ra == rb and np.array_equal(a.view(np.uint32), b.view(np.uint32))
        col = rng.uniform(-0.2, 1.3, size=3).astype(F)
        qa, qb = np.zeros(4, np.uint8), np.zeros(4, np.uint8)
        ref.ptref_make_color(col, qa)
        port.ptor_make_color(0, col, qb)
        assert np.array_equal(qa, qb)


def test_ref_struct_layout(ref):
    """ctypes mirror == the reference's own PathTraceParams (pathTracer.h:85-108) as g++ lays it out."""
    from acgpathtracing_b200.types import PathTraceParams
    assert ref.ptref_sizeof_params() == C.sizeof(PathTraceParams) == 168
    names = ["currentFrameIdx", "accumulationBuffer", "frameBuffer", "width", "height", "samplesPerPixel", "maxDepth",
             "cameraEye", "cameraU", "cameraV", "cameraW", "areaLight", "handle", "useDirectLighting",
             "useImportanceSampling"]
    for i, n in enumerate(names):
        assert ref.ptref_offsetof_params(i) == getattr(PathTraceParams, n).offset, n
    assert ref.ptref_sizeof_material() == 40 and ref.ptref_sizeof_hitgroup() == 56
