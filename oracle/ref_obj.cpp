// oracle/ref_obj.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference's own OBJ/MTL ingest (PathTracer_Optix/TinyObjWrapper.cpp on top of the
// vendored util/tiny_obj_loader.h v2.0.0) for the host, straight from the reference checkout, so
// that tests can compare the product's loader against it array by array.
#include "PathTracer_Optix/TinyObjWrapper.cpp"

#include <cstring>

extern "C" {

// Two-call protocol: first with NULL outputs to get the counts, then with buffers.
int ptrefobj_load(const char* path, uint32_t* n_vert_floats, uint32_t* n_indices, uint32_t* n_mat_ids,
                  uint32_t* n_materials, float* verts, uint32_t* indices, uint32_t* mat_ids, void* materials40)
{
    TinyObjWrapper obj;
    if (!obj.loadFile(path)) return 1;
    std::vector<float> v = obj.getVerticesFloat();            // TinyObjWrapper.cpp:200-221
    std::vector<uint32_t> ib = obj.getIndexBuffer();          // :229-244
    std::vector<uint32_t> mi = obj.getMaterialIndices();      // :177-192
    std::vector<Material> ms = obj.getMaterials();            // :138-169
    *n_vert_floats = (uint32_t)v.size();
    *n_indices = (uint32_t)ib.size();
    *n_mat_ids = (uint32_t)mi.size();
    *n_materials = (uint32_t)ms.size();
    if (verts) std::memcpy(verts, v.data(), v.size() * sizeof(float));
    if (indices) std::memcpy(indices, ib.data(), ib.size() * sizeof(uint32_t));
    if (mat_ids) std::memcpy(mat_ids, mi.data(), mi.size() * sizeof(uint32_t));
    if (materials40) std::memcpy(materials40, ms.data(), ms.size() * sizeof(Material));
    return 0;
}

}  // extern "C"
