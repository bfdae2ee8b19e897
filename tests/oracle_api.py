"""ctypes bindings of the CPU oracle (oracle/libptoracle.so) and, when present, of the reference's own
programs compiled for the host (oracle/_ref/libptref.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from acgpathtracing_b200.types import Counters, Material, PathTraceParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_vp = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_oracle():
    """make -C oracle (port always; _ref only when the reference checkout exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def load_port():
    path = os.path.join(ORACLE_DIR, "libptoracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    lib.ptor_tea4.restype = C.c_uint32
    lib.ptor_tea4.argtypes = [C.c_uint32, C.c_uint32]
    lib.ptor_rnd.restype = C.c_float
    lib.ptor_rnd.argtypes = [C.POINTER(C.c_uint32)]
    lib.ptor_lcg.restype = C.c_uint32
    lib.ptor_lcg.argtypes = [C.POINTER(C.c_uint32)]
    lib.ptor_fr_dielectric.restype = C.c_float
    lib.ptor_fr_dielectric.argtypes = [C.c_float] * 3
    lib.ptor_cosine_sample_hemisphere.argtypes = [C.c_int, C.c_float, C.c_float, _f32p]
    lib.ptor_uniform_sample_hemisphere.argtypes = [C.c_int, C.c_float, C.c_float, _f32p]
    lib.ptor_sample_ggx.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, _f32p, _f32p]
    lib.ptor_fresnel_conductor.argtypes = [C.c_float, _f32p, _f32p, _f32p]
    lib.ptor_refract.argtypes = [_f32p, _f32p, C.c_float, _f32p]
    lib.ptor_refract.restype = C.c_int
    lib.ptor_onb.argtypes = [_f32p, _f32p, _f32p]
    lib.ptor_make_color.argtypes = [C.c_int, _f32p, np.ctypeslib.ndpointer(dtype=np.uint8)]
    lib.ptor_detmath.argtypes = [C.c_int, _f32p, _f32p, C.c_uint64, C.c_float]
    lib.ptor_camera_uvw.argtypes = [_f32p, _f32p, _f32p, C.c_float, C.c_float, _f32p, _f32p, _f32p]
    lib.ptor_scene_create.restype = C.c_void_p
    lib.ptor_scene_create.argtypes = [_vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, C.c_uint32]
    lib.ptor_scene_destroy.argtypes = [_vp]
    lib.ptor_scene_set_background.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
    lib.ptor_bvh_num_nodes.restype = C.c_uint32
    lib.ptor_bvh_num_nodes.argtypes = [_vp]
    lib.ptor_bvh_export.argtypes = [_vp] * 8
    lib.ptor_scene_set_tree.argtypes = [_vp, C.c_int]
    lib.ptor_scene_set_tree.restype = C.c_int
    lib.ptor_bvh_export_sah.argtypes = [_vp, _vp]
    lib.ptor_bvh_export_sah.restype = C.c_uint32
    lib.ptor_bvh_export_ploc.argtypes = [_vp, _vp]
    lib.ptor_bvh_export_ploc.restype = C.c_uint32
    lib.ptor_trace.argtypes = [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_uint64, _vp, _vp, _vp, _vp]
    lib.ptor_render.restype = C.c_int
    lib.ptor_render.argtypes = [_vp, C.POINTER(PathTraceParams), C.c_int, C.c_int, C.c_int, _vp, _vp,
                                C.POINTER(Counters), _vp, _vp, _vp, _vp, _vp, _vp,
                                C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.ptor_hardware_threads.restype = C.c_int
    lib.ptor_scene_set_decision_probe.argtypes = [_vp, _vp]
    return lib


def load_ref():
    """The reference's own programs on the host; None when it was never built (GPU box without prebuilt _ref)."""
    path = os.path.join(ORACLE_DIR, "_ref", "libptref.so")
    if not os.path.exists(path):
        if os.path.exists("/root/reference/PathTracer_Optix/pathTracerPrograms.cu"):
            build_oracle()
        if not os.path.exists(path):
            return None
    lib = C.CDLL(path)
    lib.ptref_tea4.restype = C.c_uint32
    lib.ptref_tea4.argtypes = [C.c_uint32, C.c_uint32]
    lib.ptref_rnd.restype = C.c_float
    lib.ptref_rnd.argtypes = [C.POINTER(C.c_uint32)]
    lib.ptref_fr_dielectric.restype = C.c_float
    lib.ptref_fr_dielectric.argtypes = [C.c_float] * 3
    lib.ptref_cosine_sample_hemisphere.argtypes = [C.c_float, C.c_float, _f32p]
    lib.ptref_uniform_sample_hemisphere.argtypes = [C.c_float, C.c_float, _f32p]
    lib.ptref_sample_ggx.argtypes = [C.c_float, C.c_float, C.c_float, _f32p, _f32p]
    lib.ptref_fresnel_conductor.argtypes = [C.c_float, _f32p, _f32p, _f32p]
    lib.ptref_refract.argtypes = [_f32p, _f32p, C.c_float, _f32p]
    lib.ptref_refract.restype = C.c_int
    lib.ptref_onb.argtypes = [_f32p, _f32p, _f32p]
    lib.ptref_make_color.argtypes = [_f32p, np.ctypeslib.ndpointer(dtype=np.uint8)]
    lib.ptref_offsetof_params.restype = C.c_uint
    lib.ptref_offsetof_params.argtypes = [C.c_int]
    lib.ptref_render.restype = C.c_int
    lib.ptref_render.argtypes = [C.POINTER(PathTraceParams), _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp, C.c_uint32,
                                 _f32p, C.c_int, C.c_int, _vp, _vp, C.POINTER(Counters),
                                 C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.ptref_hardware_threads.restype = C.c_int
    return lib


class OracleScene:
    """Scene handle of the oracle port."""

    def __init__(self, lib, scene, background=(0.0, 0.0, 0.0)):
        self.lib = lib
        self.scene = scene
        self.verts = np.ascontiguousarray(scene.verts, dtype=np.float32)
        self.indices = np.ascontiguousarray(scene.indices, dtype=np.uint32)
        self.mat_ids = np.ascontiguousarray(scene.mat_ids, dtype=np.uint32)
        self.mats = scene.material_structs()
        self.h = lib.ptor_scene_create(_ptr(self.verts), self.verts.shape[0], _ptr(self.indices),
                                       self.indices.shape[0], _ptr(self.mat_ids),
                                       C.cast(self.mats, C.c_void_p), len(self.mats))
        self.background = background
        lib.ptor_scene_set_background(self.h, *background)

    def __del__(self):
        try:
            self.lib.ptor_scene_destroy(self.h)
        except Exception:
            pass

    def bvh(self):
        nT = self.indices.shape[0]
        nN = self.lib.ptor_bvh_num_nodes(self.h)
        out = dict(
            morton=np.zeros(nT, np.uint32), sorted_keys=np.zeros(nT, np.uint64), sorted_prim=np.zeros(nT, np.uint32),
            parent=np.zeros(nN, np.int32), nodes=np.zeros((nN, 16), np.uint32), tris=np.zeros((nT, 12), np.uint32),
            bounds=np.zeros(6, np.float32))
        self.lib.ptor_bvh_export(self.h, _ptr(out["morton"]), _ptr(out["sorted_keys"]), _ptr(out["sorted_prim"]),
                                 _ptr(out["parent"]), _ptr(out["nodes"]), _ptr(out["tris"]), _ptr(out["bounds"]))
        return out

    def set_tree(self, tree):
        """0 auto (the library's rule), 1 LBVH, 2 SAH, 3 PLOC; returns the tree the traversals walk."""
        return {1: "lbvh", 2: "sah", 3: "ploc"}[self.lib.ptor_scene_set_tree(self.h, tree)]

    def bvh_ploc(self):
        nodes = np.zeros((self.lib.ptor_bvh_num_nodes(self.h), 16), np.uint32)
        depth = self.lib.ptor_bvh_export_ploc(self.h, _ptr(nodes))
        return dict(nodes=nodes, depth=depth)

    def bvh_sah(self):
        nodes = np.zeros((self.lib.ptor_bvh_num_nodes(self.h), 16), np.uint32)
        depth = self.lib.ptor_bvh_export_sah(self.h, _ptr(nodes))
        return dict(nodes=nodes, depth=depth)

    def trace(self, org, dirs, tmin, tmax, use_bvh=True, want_any=False):
        org = np.ascontiguousarray(org, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = org.shape[0]
        tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, np.float32), (n,)))
        tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (n,)))
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        anyh = np.zeros(n, np.uint8) if want_any else None
        ctr = np.zeros(2, np.uint64)
        self.lib.ptor_trace(self.h, int(use_bvh), _ptr(org), _ptr(dirs), _ptr(tmin), _ptr(tmax), n, _ptr(prim),
                            _ptr(t), _ptr(anyh), _ptr(ctr))
        return prim, t, anyh, ctr

    def render(self, params: PathTraceParams, accum=None, math_mode=1, use_bvh=True, nthreads=0, probes=False,
               region=None, want_fb=True, decisions=False):
        """One launch.  Returns dict(accum (h,w,4) f32, fb (h,w,4) u8, counters, probes...); decisions=True adds
        "decisions": (h,w) u32, a hash of every discrete decision the pixel's paths took (pt_oracle.cpp decision_probe)."""
        w, h = params.width, params.height
        dec = np.zeros((h, w), np.uint32) if decisions else None
        self.lib.ptor_scene_set_decision_probe(self.h, _ptr(dec))
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        accum = np.ascontiguousarray(accum, np.float32)
        fb = np.zeros((h, w, 4), np.uint8) if want_fb else None
        ctr = Counters()
        pr = {}
        if probes:
            pr = dict(prim=np.full((h, w), -2, np.int32), t=np.zeros((h, w), np.float32),
                      seed_after_jitter=np.zeros((h, w), np.uint32), final_path_seed=np.zeros((h, w), np.uint32),
                      final_pixel_seed=np.zeros((h, w), np.uint32), primary_dir=np.zeros((h, w, 3), np.float32))
        if nthreads <= 0:
            nthreads = max(1, self.lib.ptor_hardware_threads())
        x0, x1, y0, y1 = region if region else (0, w, 0, h)
        rc = self.lib.ptor_render(self.h, C.byref(params), math_mode, int(use_bvh), nthreads, _ptr(accum), _ptr(fb),
                                  C.byref(ctr), _ptr(pr.get("prim")), _ptr(pr.get("t")),
                                  _ptr(pr.get("seed_after_jitter")), _ptr(pr.get("final_path_seed")),
                                  _ptr(pr.get("final_pixel_seed")), _ptr(pr.get("primary_dir")), x0, x1, y0, y1)
        self.lib.ptor_scene_set_decision_probe(self.h, None)
        if rc != 0:
            raise RuntimeError(f"ptor_render failed: {rc}")
        if decisions:
            pr["decisions"] = dec
        return dict(accum=accum, fb=fb, counters=ctr.asdict(), **pr)


def ref_render(lib, scene, params: PathTraceParams, accum=None, background=(0.0, 0.0, 0.0, 0.0), use_bvh=False,
               nthreads=0, region=None):
    """One launch of the reference's own raygen on the host."""
    w, h = params.width, params.height
    verts = np.ascontiguousarray(scene.verts, np.float32)
    indices = np.ascontiguousarray(scene.indices, np.uint32)
    mat_ids = np.ascontiguousarray(scene.mat_ids, np.uint32)
    mats = scene.material_structs()
    if accum is None:
        accum = np.zeros((h, w, 4), np.float32)
    accum = np.ascontiguousarray(accum, np.float32)
    fb = np.zeros((h, w, 4), np.uint8)
    ctr = Counters()
    bg = np.asarray(background, np.float32)
    if nthreads <= 0:
        nthreads = max(1, lib.ptref_hardware_threads())
    x0, x1, y0, y1 = region if region else (0, w, 0, h)
    rc = lib.ptref_render(C.byref(params), _ptr(verts), verts.shape[0], _ptr(indices), indices.shape[0],
                          _ptr(mat_ids), C.cast(mats, C.c_void_p), len(mats), bg, int(use_bvh), nthreads,
                          _ptr(accum), _ptr(fb), C.byref(ctr), x0, x1, y0, y1)
    if rc != 0:
        raise RuntimeError(f"ptref_render failed: {rc}")
    return dict(accum=accum, fb=fb, counters=ctr.asdict())


def make_params(lib, width, height, spp, max_depth, frame=0, nee=False, importance=True, scene_handle=0):
    """PathTraceParams with the reference's camera (PathTracerMain.cpp:228-233, aspect = w/h :664) and
    area light (:154-158)."""
    from acgpathtracing_b200 import scenes as S
    from acgpathtracing_b200.types import Float3
    p = PathTraceParams()
    p.currentFrameIdx = frame
    p.width, p.height, p.samplesPerPixel, p.maxDepth = width, height, spp, max_depth
    U, V, W = (np.zeros(3, np.float32) for _ in range(3))
    lib.ptor_camera_uvw(np.asarray(S.CAMERA_EYE, np.float32), np.asarray(S.CAMERA_LOOKAT, np.float32),
                        np.asarray(S.CAMERA_UP, np.float32), S.CAMERA_FOVY, np.float32(width) / np.float32(height),
                        U, V, W)
    p.cameraEye = Float3(*S.CAMERA_EYE)
    p.cameraU, p.cameraV, p.cameraW = Float3(*U), Float3(*V), Float3(*W)
    p.areaLight.corner = Float3(*S.LIGHT_CORNER)
    p.areaLight.v1 = Float3(*S.LIGHT_V1)
    p.areaLight.v2 = Float3(*S.LIGHT_V2)
    p.areaLight.normal = Float3(0.0, -1.0, 0.0)     # normalize(cross(v1, v2)), PathTracerMain.cpp:158
    p.areaLight.emission = Float3(*S.LIGHT_EMISSION)
    p.handle = scene_handle
    p.useDirectLighting = bool(nee)
    p.useImportanceSampling = bool(importance)
    return p
