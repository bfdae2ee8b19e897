"""CPU-side checks of the product's host code: the C-ABI library loads and exports every symbol that
include/ptb200.h declares, host helpers agree with the oracle / the reference's own host code, and the OBJ
ingest reproduces the reference's TinyObjWrapper arrays.  No GPU compute is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_api as O
from acgpathtracing_b200 import scenes as S
from acgpathtracing_b200.types import Material

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from acgpathtracing_b200 import api as A
    return A


def test_library_exports_every_declared_symbol(api):
    header = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 30
    raw = C.CDLL(api.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/ptb200.h but not exported"
    assert sorted(api.EXPORTED) == declared          # the Python binding covers the whole header


def test_no_gpu_fails_loudly(api):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.PtError) as e:
        api.Scene(S.cornell32())
    assert e.value.code == api.PT_ERR_CUDA


def test_struct_layout_matches_header(api):
    from acgpathtracing_b200.types import PathTraceParams, AreaLight
    assert C.sizeof(PathTraceParams) == 168 and C.sizeof(AreaLight) == 60 and C.sizeof(Material) == 40


@pytest.mark.parametrize("w,h", [(512, 512), (1920, 1080), (3840, 2160), (97, 61)])
def test_camera_frame_equals_oracle(api, port, w, h):
    U, V, W = api.camera_uvw_frame(S.CAMERA_EYE, S.CAMERA_LOOKAT, S.CAMERA_UP, S.CAMERA_FOVY, np.float32(w) / np.float32(h))
    p = O.make_params(port, w, h, 1, 1)
    for mine, ref in ((U, p.cameraU), (V, p.cameraV), (W, p.cameraW)):
        assert np.array_equal(mine.view(np.uint32), np.asarray(ref.tolist(), np.float32).view(np.uint32))
    if (w, h) == (512, 512):     # SURVEY Appendix C known answer
        assert U[0] == np.float32(-387.817566) and V[1] == np.float32(387.817566) and W[2] == 1230.0


def test_make_params_equals_oracle_params(api, port):
    a = api.make_params(1920, 1080, 128, 16, frame=3, nee=True, importance=True)
    b = O.make_params(port, 1920, 1080, 128, 16, frame=3, nee=True, importance=True)
    assert bytes(a) == bytes(b)


@pytest.mark.parametrize("w,h,n", [(1920, 1080, 8), (250, 122, 3), (8, 4, 2), (17, 9, 4), (512, 512, 1)])
def test_partition_host_logic(api, w, h, n):
    from acgpathtracing_b200 import partition as P
    cap = P.capacity(w, h, n)
    assert api.partition_capacity(w, h, n) == cap
    seen = np.zeros(w * h, np.int32)
    for r in range(n):
        idx = P.compact_pixel_index(w, h, r, n)
        assert idx.shape[0] == cap
        v = idx[idx >= 0]
        seen[v] += 1
        assert np.all(P.owner_map(w, h, n).reshape(-1)[v] == r)
    assert np.all(seen == 1)                                       # every pixel owned exactly once
    img = np.random.default_rng(1).random((h, w, 4)).astype(np.float32)
    comp = np.stack([P.pack_numpy(img, w, h, r, n) for r in range(n)])
    assert np.array_equal(P.unpack_numpy(comp, w, h, n), img)      # pack -> unpack round trip
    own = [P.owned_tiles(w, h, r, n) for r in range(n)]
    assert max(own) - min(own) <= 1                                # balanced to one tile


# ---- OBJ / MTL ingest against the reference's own TinyObjWrapper (oracle/_ref/libptrefobj.so) -----------------
def _ref_obj():
    path = os.path.join(ROOT, "oracle", "_ref", "libptrefobj.so")
    if not os.path.exists(path):
        O.load_ref()
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ptrefobj_load.restype = C.c_int
    lib.ptrefobj_load.argtypes = [C.c_char_p] + [C.POINTER(C.c_uint32)] * 4 + [C.c_void_p] * 4
    return lib


def _ref_load(lib, path):
    n = [C.c_uint32() for _ in range(4)]
    rc = lib.ptrefobj_load(path.encode(), *[C.byref(x) for x in n], None, None, None, None)
    if rc != 0:
        return None
    v = np.zeros(n[0].value, np.float32)
    i = np.zeros(n[1].value, np.uint32)
    m = np.zeros(n[2].value, np.uint32)
    mats = (Material * n[3].value)()
    rc = lib.ptrefobj_load(path.encode(), *[C.byref(x) for x in n], v.ctypes.data, i.ctypes.data, m.ctypes.data,
                           C.cast(mats, C.c_void_p))
    assert rc == 0
    return v, i, m, bytes(mats)


TRICKY_OBJ = """# comment
mtllib tricky.mtl
v 0 0 0
v 1.5 0 0
v 1.5 1 0.25
v 0 1 0
v -.5 2e0 1E-1
v +3.25 -1.0e+1 0.000123456789
v 2 2 2 0.5
v 1e-3 1.23456789012345 123456.789
v 0.1 0.2 0.3
v 4 0 0
v 5 1 0
v 4.5 2.5 0
v 3.5 2.5 0
v 3 1 0
vt 0 0
vt 1 0
vt 1 1
vn 0 0 1
f 1 2 3
usemtl red
f 1 2 3 4
f 1/1 2/2 3/3
usemtl GlassRefractive
f -4 -3 -2
f 2//1 3//1 4//1 5//1
o second
usemtl MirrorMetallic
f 10 11 12 13 14
f 14 13 12 11 10 9
g third group
usemtl nosuchmaterial
f 6 7 8
usemtl red
f 1 2
f 5 6 7 8
s off
l 1 2
f 3 2 1
"""

TRICKY_MTL = """# materials
newmtl red
Kd 0.8 0.05 0.05
Ka 1 1 1
Ks 0.5 0.5 0.5
Ns 10
newmtl GlassRefractive
Kd 1 1 1
Ni 1.5
d 0.5
newmtl MirrorMetallic   
	Kd 0.95 0.93 0.88
Pr 0.2
Pm 1
Ke 0 0 0
newmtl lightRefractiveMetallic
Ke 10 9.5 8
Kd .78 .78 .78
illum 2
"""


def _write(tmp, name, obj, mtl=None, mtlname=None):
    p = os.path.join(tmp, name + ".obj")
    with open(p, "w") as f:
        f.write(obj)
    if mtl is not None:
        with open(os.path.join(tmp, (mtlname or name) + ".mtl"), "w") as f:
            f.write(mtl)
    return p


def _compare(api, ref, path):
    r = _ref_load(ref, path)
    assert r is not None
    v, i, m, mats = r
    o = api.ObjFile(path)
    assert np.array_equal(o.verts.reshape(-1).view(np.uint32), v.view(np.uint32))
    assert np.array_equal(o.indices.reshape(-1), i)
    assert np.array_equal(o.mat_ids, m)
    assert bytes(o.materials) == mats
    return o


def test_obj_loader_matches_reference(api, tmp_path):
    ref = _ref_obj()
    if ref is None:
        pytest.skip("oracle/_ref/libptrefobj.so not built (no reference checkout)")
    tmp = str(tmp_path)
    for name, sc in (("cornell32", S.cornell32()), ("material_scene", S.material_scene(2))):
        p = _write(tmp, name, sc.obj_text, sc.mtl_text)
        o = _compare(api, ref, p)
        # and both equal the generator's arrays
        assert np.array_equal(o.verts, sc.verts) and np.array_equal(o.indices, sc.indices)
        assert np.array_equal(o.mat_ids, sc.mat_ids)
        assert bytes(o.materials) == bytes(sc.material_structs())
    o = _compare(api, ref, _write(tmp, "tricky", TRICKY_OBJ, TRICKY_MTL))
    assert o.indices.shape[0] >= 14 and (o.mat_ids == 0xFFFFFFFF).any()
    assert [m.bsdfType for m in o.materials] == [0, 2, 1, 2]
    _compare(api, ref, _write(tmp, "crlf", TRICKY_OBJ.replace("\n", "\r\n").replace("tricky.mtl", "crlf.mtl"),
                              TRICKY_MTL.replace("\n", "\r\n")))
    _compare(api, ref, _write(tmp, "nomtl", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n"))
    _compare(api, ref, _write(tmp, "missingmtl", "mtllib nothere.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl x\nf 1 2 3\n"))
    _compare(api, ref, _write(tmp, "emptymtl", "mtllib emptymtl.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "# nothing\nKd 1 0 0\n"))
    # concave and random polygons (ear clipping) + random quads (diagonal choice)
    rng = np.random.default_rng(3)
    lines = []
    nv = 0
    for k in range(60):
        n = int(rng.integers(4, 9))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        rad = rng.uniform(0.3, 1.0, n)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.normal(0, 0.05, n)], 1) + rng.normal(0, 3, 3)
        rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pts = pts @ rot
        lines += ["v %.7g %.7g %.7g" % tuple(p) for p in pts]
        lines.append("f " + " ".join(str(nv + 1 + j) for j in range(n)))
        nv += n
    _compare(api, ref, _write(tmp, "polys", "\n".join(lines) + "\n"))


def test_obj_loader_errors(api, tmp_path):
    with pytest.raises(api.PtError) as e:
        api.ObjFile(str(tmp_path / "does_not_exist.obj"))
    assert e.value.code == api.PT_ERR_IO
    p = _write(str(tmp_path), "zero", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n")
    with pytest.raises(api.PtError) as e:
        api.ObjFile(p)
    assert e.value.code == api.PT_ERR_IO
    # (tinyobjloader v2.0.0 specifies "return false" for a zero index, tiny_obj_loader.h:830-841,2795-2806; the
    #  reference build itself crashes on this input, so it is not executed here.)


# ---- image output + checkpoint (host only; sutil::saveImage convention, sutil/sutil.cpp:543-570) -----------------------
def test_ppm_writer_flips_rows_and_drops_alpha(api, tmp_path):
    w, h = 5, 3
    img = np.arange(w * h * 4, dtype=np.uint8).reshape(h, w, 4)          # row 0 = bottom of the picture
    path = str(tmp_path / "a.ppm")
    api._check(api.lib.pt_image_write_ppm(path.encode(), img.ctypes.data, w, h))
    raw = open(path, "rb").read()
    head = f"P6\n{w} {h}\n255\n".encode()
    assert raw.startswith(head) and len(raw) == len(head) + 3 * w * h
    pix = np.frombuffer(raw[len(head):], np.uint8).reshape(h, w, 3)
    assert np.array_equal(pix, img[::-1, :, :3])                          # written top-down
    with pytest.raises(api.PtError) as e:
        api._check(api.lib.pt_image_write_ppm(str(tmp_path / "no_dir" / "a.ppm").encode(), img.ctypes.data, w, h))
    assert e.value.code == api.PT_ERR_IO
    with pytest.raises(api.PtError):
        api._check(api.lib.pt_image_write_ppm(path.encode(), None, w, h))


def test_png_writer_and_extension_dispatch(api, tmp_path):
    """The ".png" branch of sutil::saveImage (sutil/sutil.cpp:622-638): RGBA, flipped vertically; decoded here with PIL.  Sizes
    include one whose raw data spans several 65535-byte stored blocks."""
    from PIL import Image
    rng = np.random.default_rng(11)
    for w, h in ((5, 3), (1, 1), (301, 257)):
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)             # row 0 = bottom of the picture
        path = str(tmp_path / f"a{w}.png")
        api._check(api.lib.pt_image_write_png(path.encode(), img.ctypes.data, w, h))
        got = Image.open(path)
        assert got.mode == "RGBA" and got.size == (w, h)
        assert np.array_equal(np.asarray(got), img[::-1])
        via = str(tmp_path / f"b{w}.PNG")
        api._check(api.lib.pt_image_write(via.encode(), img.ctypes.data, w, h))
        assert open(via, "rb").read() == open(path, "rb").read()
    ppm = str(tmp_path / "c.ppm")
    api._check(api.lib.pt_image_write(ppm.encode(), img.ctypes.data, w, h))
    assert open(ppm, "rb").read().startswith(b"P6")
    for bad in ("x.jpg", "a.exr", "noextension"):
        with pytest.raises(api.PtError) as e:
            api._check(api.lib.pt_image_write(str(tmp_path / bad).encode(), img.ctypes.data, w, h))
        assert e.value.code == api.PT_ERR_INVALID


def test_accumulation_checkpoint_round_trip(api, tmp_path):
    w, h = 7, 4
    acc = np.random.default_rng(3).random((h, w, 4)).astype(np.float32)
    path = str(tmp_path / "acc.bin")
    api._check(api.lib.pt_accum_save(path.encode(), acc.ctypes.data, w, h, 5))
    back = np.zeros_like(acc)
    nxt = C.c_uint32()
    api._check(api.lib.pt_accum_load(path.encode(), back.ctypes.data, w, h, C.byref(nxt)))
    assert nxt.value == 5 and np.array_equal(back.view(np.uint32), acc.view(np.uint32))
    with pytest.raises(api.PtError) as e:                                 # wrong size
        api._check(api.lib.pt_accum_load(path.encode(), back.ctypes.data, w + 1, h, C.byref(nxt)))
    assert e.value.code == api.PT_ERR_INVALID
    open(path, "wb").write(b"garbage")
    with pytest.raises(api.PtError) as e:
        api._check(api.lib.pt_accum_load(path.encode(), back.ctypes.data, w, h, C.byref(nxt)))
    assert e.value.code == api.PT_ERR_IO


def test_headless_driver_is_built_and_explains_itself():
    exe = os.path.join(ROOT, "acgpathtracing_b200", "ptb_render")
    assert os.path.exists(exe), "build with make -C acgpathtracing_b200/csrc"
    import subprocess
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: ptb_render --obj" in r.stderr
