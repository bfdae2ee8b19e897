"""ctypes mirrors of the reference's launch contract.

Byte-identical to /root/reference/PathTracer_Optix/pathTracer.h:77-108 (AreaLight, PathTraceParams) and
TinyObjWrapper.h:27-40 (BSDFType, Material); the layout is asserted against the reference's own headers
compiled on the host in tests/test_oracle_vs_ref.py.
"""
import ctypes as C

BSDF_DIFFUSE = 0      # TinyObjWrapper.h:27-31
BSDF_METALLIC = 1
BSDF_REFRACTION = 2


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(float(x), float(y), float(z))

    def tolist(self):
        return [self.x, self.y, self.z]


class AreaLight(C.Structure):          # pathTracer.h:77-83 (60 bytes)
    _fields_ = [("corner", Float3), ("v1", Float3), ("v2", Float3), ("normal", Float3), ("emission", Float3)]


class PathTraceParams(C.Structure):    # pathTracer.h:85-108 (168 bytes, align 8)
    _fields_ = [
        ("currentFrameIdx", C.c_uint32),
        ("accumulationBuffer", C.c_void_p),   # float4*  (device pointer for the CUDA library)
        ("frameBuffer", C.c_void_p),          # uchar4*  (may be NULL: headless extension)
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("samplesPerPixel", C.c_uint32),
        ("maxDepth", C.c_uint32),
        ("cameraEye", Float3),
        ("cameraU", Float3),
        ("cameraV", Float3),
        ("cameraW", Float3),
        ("areaLight", AreaLight),
        ("handle", C.c_uint64),               # OptixTraversableHandle in the reference; our scene handle
        ("useDirectLighting", C.c_bool),
        ("useImportanceSampling", C.c_bool),
    ]


class Material(C.Structure):           # TinyObjWrapper.h:33-40 (40 bytes)
    _fields_ = [
        ("diffuse", Float3),
        ("emission", Float3),
        ("roughness", C.c_float),
        ("metallic", C.c_float),
        ("ior", C.c_float),
        ("bsdfType", C.c_int32),
    ]


class Counters(C.Structure):
    """Work counters for the Grays/s metric and the algorithmic-work roofline."""
    _fields_ = [
        ("n_samples", C.c_uint64),
        ("n_radiance", C.c_uint64),    # radiance segments traced (closest-hit traversals)
        ("n_shadow", C.c_uint64),      # shadow rays traced (any-hit traversals)
        ("n_node", C.c_uint64),        # BVH nodes fetched (2 slab tests each)
        ("n_tri", C.c_uint64),         # ray/triangle tests
        ("n_hit", C.c_uint64 * 3),     # closest hits by BSDF type
        ("n_miss", C.c_uint64),
    ]

    def asdict(self):
        return {
            "n_samples": self.n_samples, "n_radiance": self.n_radiance, "n_shadow": self.n_shadow,
            "n_node": self.n_node, "n_tri": self.n_tri, "n_hit": list(self.n_hit), "n_miss": self.n_miss,
        }


assert C.sizeof(PathTraceParams) == 168
assert C.sizeof(AreaLight) == 60
assert C.sizeof(Material) == 40
assert PathTraceParams.accumulationBuffer.offset == 8
assert PathTraceParams.width.offset == 24
assert PathTraceParams.cameraEye.offset == 40
assert PathTraceParams.areaLight.offset == 88
assert PathTraceParams.handle.offset == 152
assert PathTraceParams.useDirectLighting.offset == 160
assert PathTraceParams.useImportanceSampling.offset == 161
