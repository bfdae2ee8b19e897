"""ctypes binding of the C-ABI library (include/ptb200.h -> acgpathtracing_b200/libptb200.so).

Host-side mirror of what PathTracerMain.cpp does around the hot path: build a scene from the four
TinyObjWrapper arrays (:260-398, :544-608), fill a ``PathTraceParams`` and launch (:184-210).  There is no
CPU fallback: importing this module raises when the CUDA library has not been built, and every call that
needs the GPU raises ``PtError`` when no sm_100 device is usable.
"""
import ctypes as C
import os

import numpy as np

from .types import Counters, Float3, Material, PathTraceParams
from . import scenes as _scenes

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libptb200.so")

PT_OK, PT_ERR_INVALID, PT_ERR_CUDA, PT_ERR_IO = 0, 1, 2, 3


class PtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ptb200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C acgpathtracing_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    PP = C.POINTER(PathTraceParams)
    sig = {
        "pt_last_error": (C.c_char_p, []),
        "pt_scene_create": (i32, [vp, u32, vp, u32, vp, vp, u32, C.POINTER(vp)]),
        "pt_scene_destroy": (None, [vp]),
        "pt_scene_set_background": (i32, [vp, C.c_float, C.c_float, C.c_float]),
        "pt_scene_handle": (u64, [vp]),
        "pt_scene_info": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_float)]),
        "pt_launch": (i32, [PP, vp]),
        "pt_launch_partition": (i32, [PP, u32, u32, vp]),
        "pt_launch_raw": (i32, [PP, u32, u32, vp]),
        "pt_fold_frames": (i32, [vp, vp, C.POINTER(vp), u32, u32, u32, u32, vp]),
        "pt_partition_capacity": (u64, [u32, u32, u32]),
        "pt_partition_pack": (i32, [vp, vp, vp, vp, u32, u32, u32, u32, vp]),
        "pt_partition_unpack": (i32, [vp, vp, vp, vp, u32, u32, u32, vp]),
        "pt_launch_host": (i32, [PP, C.POINTER(u64), C.POINTER(u64)]),
        "pt_scene_enable_counters": (i32, [vp, i32]),
        "pt_scene_get_counters": (i32, [vp, C.POINTER(Counters), i32]),
        "pt_scene_get_sched_stats": (i32, [vp, C.POINTER(u64), i32]),
        "pt_scene_set_pipeline": (i32, [vp, i32]),
        "pt_scene_set_traversal": (i32, [vp, i32]),
        "pt_scene_set_tune": (i32, [vp, C.c_char_p]),
        "pt_scene_check_error": (i32, [vp]),
        "pt_multi_plan": (i32, [u32, u32, vp, u32, C.POINTER(u32)]),
        "pt_multi_slice": (i32, [u32, u32, u32, u32, C.POINTER(u64), C.POINTER(u64)]),
        "pt_group_create_local": (i32, [C.POINTER(C.c_int), u32, i32, C.POINTER(vp)]),
        "pt_group_unique_id": (i32, [vp]),
        "pt_group_create_rank": (i32, [vp, u32, u32, C.POINTER(vp)]),
        "pt_group_destroy": (None, [vp]),
        "pt_group_info": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(i32)]),
        "pt_group_scene_create": (i32, [vp, vp, u32, vp, u32, vp, vp, u32, C.POINTER(vp)]),
        "pt_group_scene": (vp, [vp, u32]),
        "pt_launch_multi": (i32, [vp, PP, u32, vp]),
        "pt_launch_multi_host": (i32, [vp, PP, u32, C.POINTER(u64), C.POINTER(u64)]),
        "pt_group_last_timing": (i32, [vp, u32, C.POINTER(C.c_float), C.POINTER(u32)]),
        "pt_scene_set_tree": (i32, [vp, i32]),
        "pt_scene_tree_in_use": (i32, [vp]),
        "pt_bvh_export_sah": (i32, [vp, vp, C.POINTER(u32)]),
        "pt_bvh_export_ploc": (i32, [vp, vp, C.POINTER(u32), C.POINTER(C.c_float)]),
        "pt_bvh_export_quant": (i32, [vp, vp, vp, vp]),
        "pt_bvh_export_wide": (i32, [vp, vp, u32, vp, vp, vp]),
        "pt_sah_build_host": (i32, [vp, vp, u32, vp, C.POINTER(u32)]),
        "pt_scene_last_traversal": (i32, [vp]),
        "pt_scene_last_launch_info": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "pt_bvh_export": (i32, [vp] * 8),
        "pt_trace_rays": (i32, [vp, vp, vp, vp, vp, u64, vp, vp, vp, vp]),
        "pt_debug_primary": (i32, [PP, vp, vp, vp, vp]),
        "pt_debug_final_seeds": (i32, [PP, vp, vp]),
        "pt_camera_uvw_frame": (None, [vp, vp, vp, C.c_float, C.c_float, vp, vp, vp]),
        "pt_obj_load": (i32, [C.c_char_p, C.POINTER(vp)]),
        "pt_obj_free": (None, [vp]),
        "pt_obj_counts": (i32, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "pt_obj_copy": (i32, [vp, vp, vp, vp, vp]),
        "pt_scene_create_from_obj": (i32, [vp, C.POINTER(vp)]),
        "pt_image_write_ppm": (i32, [C.c_char_p, vp, u32, u32]),
        "pt_image_write_png": (i32, [C.c_char_p, vp, u32, u32]),
        "pt_image_write": (i32, [C.c_char_p, vp, u32, u32]),
        "pt_accum_save": (i32, [C.c_char_p, vp, u32, u32, u32]),
        "pt_accum_load": (i32, [C.c_char_p, vp, u32, u32, C.POINTER(u32)]),
        "pt_device_info": (i32, [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.c_char_p]),
        "pt_bench_fp32_peak": (i32, [C.POINTER(C.c_double)]),
        "pt_bench_l2_bandwidth": (i32, [C.POINTER(C.c_double)]),
        "pt_bench_hbm_bandwidth": (i32, [C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()


def _check(rc):
    if rc != PT_OK:
        raise PtError(rc, lib.pt_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sah_build_host(tri_lo, tri_hi, sorted_prim):
    """The library's host SAH builder (sah.cu) on its own: per-primitive boxes (n, 3) + leaf order -> (nodes (n-1, 16) u32, depth)."""
    n = len(sorted_prim)
    boxes = np.zeros((n, 8), np.float32)
    boxes[:, 0:3] = tri_lo
    boxes[:, 4:7] = tri_hi
    prim = np.ascontiguousarray(sorted_prim, np.uint32)
    nodes = np.zeros((n - 1, 16), np.uint32)
    d = C.c_uint32()
    _check(lib.pt_sah_build_host(_ptr(boxes), _ptr(prim), n, _ptr(nodes), C.byref(d)))
    return nodes, d.value


def camera_uvw_frame(eye, lookat, up, fov_y_deg, aspect):
    """sutil::Camera::UVWFrame (sutil/Camera.cpp:33-45)."""
    e, l, u = (np.asarray(x, np.float32) for x in (eye, lookat, up))
    U, V, W = (np.zeros(3, np.float32) for _ in range(3))
    lib.pt_camera_uvw_frame(_ptr(e), _ptr(l), _ptr(u), np.float32(fov_y_deg), np.float32(aspect), _ptr(U), _ptr(V), _ptr(W))
    return U, V, W


def make_params(width, height, spp, max_depth, frame=0, nee=False, importance=True, scene=None):
    """PathTraceParams with the reference's hard-coded camera (PathTracerMain.cpp:228-233, aspect = w/h :664)
    and area light (:154-158)."""
    p = PathTraceParams()
    p.currentFrameIdx = frame
    p.width, p.height, p.samplesPerPixel, p.maxDepth = width, height, spp, max_depth
    U, V, W = camera_uvw_frame(_scenes.CAMERA_EYE, _scenes.CAMERA_LOOKAT, _scenes.CAMERA_UP, _scenes.CAMERA_FOVY,
                               np.float32(width) / np.float32(height))
    p.cameraEye = Float3(*_scenes.CAMERA_EYE)
    p.cameraU, p.cameraV, p.cameraW = Float3(*U), Float3(*V), Float3(*W)
    p.areaLight.corner = Float3(*_scenes.LIGHT_CORNER)
    p.areaLight.v1 = Float3(*_scenes.LIGHT_V1)
    p.areaLight.v2 = Float3(*_scenes.LIGHT_V2)
    p.areaLight.normal = Float3(0.0, -1.0, 0.0)       # normalize(cross(v1, v2)), PathTracerMain.cpp:158
    p.areaLight.emission = Float3(*_scenes.LIGHT_EMISSION)
    p.handle = scene.handle if scene is not None else 0
    p.useDirectLighting = bool(nee)
    p.useImportanceSampling = bool(importance)
    return p


class ObjFile:
    """TinyObjWrapper-equivalent ingest (TinyObjWrapper.cpp:38-244): the four arrays of an OBJ/MTL pair."""

    def __init__(self, path):
        h = C.c_void_p()
        _check(lib.pt_obj_load(os.fsencode(path), C.byref(h)))
        self._h = h
        nv, ni, nm, nmat = (C.c_uint32() for _ in range(4))
        _check(lib.pt_obj_counts(h, C.byref(nv), C.byref(ni), C.byref(nm), C.byref(nmat)))
        self.verts = np.zeros((nv.value // 4, 4), np.float32)
        self.indices = np.zeros((ni.value // 3, 3), np.uint32)
        self.mat_ids = np.zeros(nm.value, np.uint32)
        self.materials = (Material * nmat.value)()
        _check(lib.pt_obj_copy(h, _ptr(self.verts), _ptr(self.indices), _ptr(self.mat_ids),
                               C.cast(self.materials, C.c_void_p)))

    def close(self):
        if self._h:
            lib.pt_obj_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Device scene + LBVH (replaces buildTheAccelarationStructure + createShaderBindingTable)."""

    def __init__(self, arrays=None, obj: ObjFile = None, background=None, _borrowed=None):
        h = C.c_void_p()
        self._owned = _borrowed is None
        if _borrowed is not None:                     # a replica owned by a Group (pt_group_scene)
            h = C.c_void_p(_borrowed)
            self.n_tri = None
        elif obj is not None:
            _check(lib.pt_scene_create_from_obj(obj._h, C.byref(h)))
            self.n_tri = int(obj.indices.shape[0])
        else:
            self._verts = np.ascontiguousarray(arrays.verts, np.float32)
            self._indices = np.ascontiguousarray(arrays.indices, np.uint32)
            self._mat_ids = np.ascontiguousarray(arrays.mat_ids, np.uint32)
            self._mats = arrays.material_structs() if hasattr(arrays, "material_structs") else arrays.materials
            _check(lib.pt_scene_create(_ptr(self._verts), self._verts.shape[0], _ptr(self._indices),
                                       self._indices.shape[0], _ptr(self._mat_ids),
                                       C.cast(self._mats, C.c_void_p), len(self._mats), C.byref(h)))
            self.n_tri = int(self._indices.shape[0])
        self._h = h
        if background is not None:
            self.set_background(*background)

    @property
    def handle(self):
        return lib.pt_scene_handle(self._h)

    def destroy(self):
        if self._h and self._owned:
            lib.pt_scene_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def set_background(self, r, g, b):
        _check(lib.pt_scene_set_background(self._h, r, g, b))

    def info(self):
        nt, nn, ms = C.c_uint32(), C.c_uint32(), C.c_float()
        _check(lib.pt_scene_info(self._h, C.byref(nt), C.byref(nn), C.byref(ms)))
        return dict(n_tris=nt.value, n_nodes=nn.value, build_ms=ms.value)

    def enable_counters(self, on=True):
        """False / 0 off; True / 1 the specified work (equals the oracle's counts); 2 the executed work of an uncounted launch."""
        _check(lib.pt_scene_enable_counters(self._h, int(on)))

    def counters(self, reset=True):
        c = Counters()
        _check(lib.pt_scene_get_counters(self._h, C.byref(c), int(reset)))
        return c.asdict()

    def sched_stats(self, reset=True):
        """Batch occupancy of counted wavefront launches: per stage batches, mean lanes per batch; NODE steps."""
        a = (C.c_uint64 * 16)()
        _check(lib.pt_scene_get_sched_stats(self._h, a, int(reset)))
        names = ("node", "leaf", "shade", "end")
        out = {n: {"batches": int(a[i]), "lanes_per_batch": (a[4 + i] / a[i]) if a[i] else 0.0} for i, n in enumerate(names)}
        out["node_steps"] = int(a[8])
        out["lanes_per_node_step"] = (a[9] / a[8]) if a[8] else 0.0
        out["idle_votes"] = int(a[10])
        return out

    def set_pipeline(self, pipeline):
        _check(lib.pt_scene_set_pipeline(self._h, pipeline))

    def set_tree(self, tree):
        """0 auto (SAH for scenes of 2..8192 triangles, PLOC above), 1 LBVH, 2 SAH, 3 PLOC (pt_scene_set_tree)."""
        _check(lib.pt_scene_set_tree(self._h, tree))

    def tree_in_use(self):
        return {1: "lbvh", 2: "sah", 3: "ploc"}.get(lib.pt_scene_tree_in_use(self._h), "?")

    def bvh_export_ploc(self):
        """dict(nodes (n_nodes, 16) u32, depth, build_ms) of the GPU-built PLOC tree."""
        nodes = np.zeros((self.info()["n_nodes"], 16), np.uint32)
        d, ms = C.c_uint32(), C.c_float()
        _check(lib.pt_bvh_export_ploc(self._h, _ptr(nodes), C.byref(d), C.byref(ms)))
        return dict(nodes=nodes, depth=d.value, build_ms=ms.value)

    def bvh_export_sah(self):
        """dict(nodes (n_nodes, 16) u32, depth) of the sweep-SAH tree."""
        nodes = np.zeros((self.info()["n_nodes"], 16), np.uint32)
        d = C.c_uint32()
        _check(lib.pt_bvh_export_sah(self._h, _ptr(nodes), C.byref(d)))
        return dict(nodes=nodes, depth=d.value)

    def bvh_export_quant(self):
        """Quantised nodes of the tree in use: dict(q (n_nodes, 8) u32, lo (3,), step (3,))."""
        q = np.zeros((self.info()["n_nodes"], 8), np.uint32)
        lo, step = np.zeros(3, np.float32), np.zeros(3, np.float32)
        _check(lib.pt_bvh_export_quant(self._h, _ptr(q), _ptr(lo), _ptr(step)))
        return dict(q=q, lo=lo, step=step)

    def bvh_export_wide(self):
        """4-wide quantised nodes of the tree in use: dict(w (n_wide, 16) u32, depth, build_ms)."""
        n, depth, ms = C.c_uint32(0), C.c_uint32(0), C.c_float(0)
        _check(lib.pt_bvh_export_wide(self._h, None, 0, C.byref(n), C.byref(depth), C.byref(ms)))
        w = np.zeros((n.value, 16), np.uint32)
        _check(lib.pt_bvh_export_wide(self._h, _ptr(w), n.value, C.byref(n), C.byref(depth), C.byref(ms)))
        return dict(w=w, depth=depth.value, build_ms=ms.value)

    def set_traversal(self, traversal):
        """0 auto, 1 exact nodes (staged in shared memory when the scene fits), 2 exact nodes from global memory,
        3 quantised nodes from global memory, 4 4-wide quantised nodes from global memory (pt_scene_set_traversal)."""
        _check(lib.pt_scene_set_traversal(self._h, traversal))

    def set_tune(self, spec=""):
        """Scheduling constants of the render kernel for this scene, "key=value,..." ("" = defaults)."""
        _check(lib.pt_scene_set_tune(self._h, spec.encode()))

    def check_error(self):
        _check(lib.pt_scene_check_error(self._h))

    def last_launch_info(self):
        g, b, s, k = (C.c_uint32() for _ in range(4))
        _check(lib.pt_scene_last_launch_info(self._h, C.byref(g), C.byref(b), C.byref(s), C.byref(k)))
        return dict(grid=g.value, block=b.value, smem_bytes=s.value, kernels=k.value,
                    traversal={0: "binary", 1: "binary-quant", 2: "wide4-quant"}.get(lib.pt_scene_last_traversal(self._h), "?"))

    def bvh_export(self):
        info = self.info()
        nT, nN = info["n_tris"], info["n_nodes"]
        out = dict(morton=np.zeros(nT, np.uint32), sorted_keys=np.zeros(nT, np.uint64),
                   sorted_prim=np.zeros(nT, np.uint32), parent=np.zeros(max(nN, 1) if nT else 0, np.int32),
                   nodes=np.zeros((nN, 16), np.uint32), tris=np.zeros((nT, 12), np.uint32), bounds=np.zeros(6, np.float32))
        _check(lib.pt_bvh_export(self._h, _ptr(out["morton"]), _ptr(out["sorted_keys"]), _ptr(out["sorted_prim"]),
                                 _ptr(out["parent"]), _ptr(out["nodes"]), _ptr(out["tris"]), _ptr(out["bounds"])))
        return out

    def trace(self, org, dirs, tmin, tmax, want_any=False):
        org = np.ascontiguousarray(org, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        n = org.shape[0]
        tmin = np.ascontiguousarray(np.broadcast_to(np.asarray(tmin, np.float32), (n,)))
        tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (n,)))
        prim = np.zeros(n, np.int32)
        t = np.zeros(n, np.float32)
        anyh = np.zeros(n, np.uint8) if want_any else None
        ctr = np.zeros(2, np.uint64)
        _check(lib.pt_trace_rays(self._h, _ptr(org), _ptr(dirs), _ptr(tmin), _ptr(tmax), n, _ptr(prim), _ptr(t),
                                 _ptr(anyh), _ptr(ctr)))
        return prim, t, anyh, ctr

    # ---- launches -------------------------------------------------------------------------------------
    def launch_host(self, params: PathTraceParams, accum=None, want_fb=True):
        """One launch with HOST buffers (pt_launch_host).  Returns dict(accum, fb, h2d_bytes, d2h_bytes)."""
        w, h = params.width, params.height
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        accum = np.ascontiguousarray(accum, np.float32)
        fb = np.zeros((h, w, 4), np.uint8) if want_fb else None
        p = PathTraceParams.from_buffer_copy(params)
        p.handle = self.handle
        p.accumulationBuffer = accum.ctypes.data
        p.frameBuffer = fb.ctypes.data if fb is not None else None
        up, down = C.c_uint64(), C.c_uint64()
        _check(lib.pt_launch_host(C.byref(p), C.byref(up), C.byref(down)))
        return dict(accum=accum, fb=fb, h2d_bytes=up.value, d2h_bytes=down.value)

    def debug_primary(self, params: PathTraceParams):
        w, h = params.width, params.height
        p = PathTraceParams.from_buffer_copy(params)
        p.handle = self.handle
        out = dict(seed_after_jitter=np.zeros((h, w), np.uint32), primary_dir=np.zeros((h, w, 3), np.float32),
                   prim=np.zeros((h, w), np.int32), t=np.zeros((h, w), np.float32))
        _check(lib.pt_debug_primary(C.byref(p), _ptr(out["seed_after_jitter"]), _ptr(out["primary_dir"]),
                                    _ptr(out["prim"]), _ptr(out["t"])))
        return out

    def debug_final_seeds(self, params: PathTraceParams):
        w, h = params.width, params.height
        p = PathTraceParams.from_buffer_copy(params)
        p.handle = self.handle
        out = dict(final_path_seed=np.zeros((h, w), np.uint32), final_pixel_seed=np.zeros((h, w), np.uint32))
        _check(lib.pt_debug_final_seeds(C.byref(p), _ptr(out["final_path_seed"]), _ptr(out["final_pixel_seed"])))
        return out


def launch(params: PathTraceParams, stream=None):
    """pt_launch: device buffers in params, asynchronous on `stream` (a raw cudaStream_t value or None)."""
    _check(lib.pt_launch(C.byref(params), C.c_void_p(stream) if stream else None))


def launch_partition(params: PathTraceParams, rank, n_ranks, stream=None):
    _check(lib.pt_launch_partition(C.byref(params), rank, n_ranks, C.c_void_p(stream) if stream else None))


def launch_raw(params, rank=0, n_ranks=1, stream=None):
    """This launch's own per-pixel mean into params.accumulationBuffer (no fold, no frame buffer)."""
    _check(lib.pt_launch_raw(C.byref(params), rank, n_ranks, C.c_void_p(stream) if stream else None))


def fold_frames(accum, frame_out, frame_means, first_frame_idx, width, height, stream=None):
    """accum / frame_out / frame_means[k]: raw device addresses.  Running average over the frames, in order."""
    arr = (C.c_void_p * len(frame_means))(*[C.c_void_p(int(p)) for p in frame_means])
    _check(lib.pt_fold_frames(C.c_void_p(int(accum)), C.c_void_p(int(frame_out)) if frame_out else None, arr,
                              first_frame_idx, len(frame_means), width, height, C.c_void_p(stream) if stream else None))


def partition_capacity(width, height, n_ranks):
    return int(lib.pt_partition_capacity(width, height, n_ranks))


def partition_pack(accum_full, frame_full, accum_compact, frame_compact, width, height, rank, n_ranks, stream=None):
    """All pointer arguments are raw device addresses (ints); frame pointers may be 0/None."""
    _check(lib.pt_partition_pack(accum_full, frame_full or None, accum_compact, frame_compact or None, width, height,
                                 rank, n_ranks, C.c_void_p(stream) if stream else None))


def partition_unpack(accum_compact_all, frame_compact_all, accum_full, frame_full, width, height, n_ranks, stream=None):
    _check(lib.pt_partition_unpack(accum_compact_all, frame_compact_all or None, accum_full, frame_full or None, width,
                                   height, n_ranks, C.c_void_p(stream) if stream else None))


# ---- multi-GPU (pt_group / pt_launch_multi) -----------------------------------------------------------------------------
class MultiJob(C.Structure):
    _fields_ = [("frame", C.c_uint32), ("rank", C.c_uint32), ("first_class", C.c_uint32), ("n_classes", C.c_uint32)]


def multi_plan(n_launches, n_ranks):
    """pt_multi_plan (host only): [(frame, rank, first_class, n_classes)] in exchange order."""
    n = C.c_uint32()
    _check(lib.pt_multi_plan(n_launches, n_ranks, None, 0, C.byref(n)))
    jobs = (MultiJob * max(n.value, 1))()
    _check(lib.pt_multi_plan(n_launches, n_ranks, C.cast(jobs, C.c_void_p), n.value, C.byref(n)))
    return [(j.frame, j.rank, j.first_class, j.n_classes) for j in jobs[:n.value]]


def multi_slice(width, height, rank, n_ranks):
    """pt_multi_slice (host only): (first pixel, pixel count) of the slice `rank` folds."""
    a, b = C.c_uint64(), C.c_uint64()
    _check(lib.pt_multi_slice(width, height, rank, n_ranks, C.byref(a), C.byref(b)))
    return a.value, b.value


PT_TRANSPORT_PEER, PT_TRANSPORT_NCCL = 0, 1
PT_GROUP_ID_BYTES = 128


def group_unique_id():
    buf = C.create_string_buffer(PT_GROUP_ID_BYTES)
    _check(lib.pt_group_unique_id(buf))
    return buf.raw


class Group:
    """pt_group: several GPUs rendering one image.  Group.local(devices) = one process drives them all;
    Group.rank(id_bytes, rank, world) = one process per GPU (collective; NCCL)."""

    def __init__(self, handle):
        self._h = handle
        self.scene = None

    @classmethod
    def local(cls, devices, transport=PT_TRANSPORT_PEER):
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib.pt_group_create_local(arr, len(devices), transport, C.byref(h)))
        return cls(h)

    @classmethod
    def rank(cls, id_bytes, rank, world):
        h = C.c_void_p()
        _check(lib.pt_group_create_rank(C.c_char_p(id_bytes), rank, world, C.byref(h)))
        return cls(h)

    def info(self):
        n, nl, fr, tr = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_int()
        _check(lib.pt_group_info(self._h, C.byref(n), C.byref(nl), C.byref(fr), C.byref(tr)))
        return dict(n_ranks=n.value, n_local=nl.value, first_rank=fr.value, transport={0: "peer", 1: "nccl"}[tr.value])

    def create_scene(self, arrays):
        """Replicates the scene on every local device; returns the first replica (its handle goes into the params)."""
        verts = np.ascontiguousarray(arrays.verts, np.float32)
        indices = np.ascontiguousarray(arrays.indices, np.uint32)
        mat_ids = np.ascontiguousarray(arrays.mat_ids, np.uint32)
        mats = arrays.material_structs() if hasattr(arrays, "material_structs") else arrays.materials
        h = C.c_void_p()
        _check(lib.pt_group_scene_create(self._h, _ptr(verts), verts.shape[0], _ptr(indices), indices.shape[0], _ptr(mat_ids),
                                         C.cast(mats, C.c_void_p), len(mats), C.byref(h)))
        self.scene = Scene(_borrowed=h.value)
        self.scene.n_tri = int(indices.shape[0])
        return self.scene

    def replica(self, local_index):
        h = lib.pt_group_scene(self._h, local_index)
        if not h:
            raise PtError(PT_ERR_INVALID, "no such replica")
        return Scene(_borrowed=h)

    def launch(self, params: PathTraceParams, n_launches, stream=None):
        """pt_launch_multi: device buffers, stream-ordered."""
        _check(lib.pt_launch_multi(self._h, C.byref(params), n_launches, C.c_void_p(stream) if stream else None))

    def launch_host(self, params: PathTraceParams, n_launches):
        """pt_launch_multi_host: host buffers, synchronous.  Returns (h2d_bytes, d2h_bytes) moved by this process."""
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib.pt_launch_multi_host(self._h, C.byref(params), n_launches, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_kernels(self, local_index=0):
        return self.last_timing(local_index)["kernels"]

    def last_timing(self, local_index=0):
        ms = (C.c_float * 4)()
        n = C.c_uint32()
        _check(lib.pt_group_last_timing(self._h, local_index, ms, C.byref(n)))
        return dict(render_ms=ms[0], exchange_fold_ms=ms[1], gather_ms=ms[2], total_ms=ms[3], kernels=n.value)

    def destroy(self):
        if self._h:
            lib.pt_group_destroy(self._h)
            self._h = None
            self.scene = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def device_info():
    sm, l2, khz = C.c_int(), C.c_int(), C.c_int()
    name = C.create_string_buffer(64)
    _check(lib.pt_device_info(C.byref(sm), C.byref(l2), C.byref(khz), name))
    return dict(sm_count=sm.value, l2_bytes=l2.value, clock_khz=khz.value, name=name.value.decode())


def bench_fp32_peak():
    v = C.c_double()
    _check(lib.pt_bench_fp32_peak(C.byref(v)))
    return v.value


def bench_l2_bandwidth():
    v = C.c_double()
    _check(lib.pt_bench_l2_bandwidth(C.byref(v)))
    return v.value


def bench_hbm_bandwidth():
    v = C.c_double()
    _check(lib.pt_bench_hbm_bandwidth(C.byref(v)))
    return v.value
