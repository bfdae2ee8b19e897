// ptb_render -- headless counterpart of the reference's main() (PathTracer_Optix/PathTracerMain.cpp:648-745) on top
// of the C ABI (include/ptb200.h).  Same defaults (512x512, 128 spp per launch, maxDepth 4, direct lighting and
// importance sampling off, camera of initCamera :228-233, area light of initializeTheLaunch :154-158), the key
// toggles of keyCallback (:100-141) as flags, the per-frame time print (:726) and the totals (:738-740); instead of
// a GLFW window the last frame is written as a PPM (sutil::saveImage convention) and the accumulation buffer can be
// saved / resumed.
//
//   ptb_render --obj scene.obj [--width W --height H --spp N --depth D --direct --importance --frames F]
//              [--background r g b] [--out image.ppm] [--accum-out file] [--accum-in file] [--pipeline 0|1|2]
//              [--gpus N | --devices 0,1,...] [--transport peer|nccl]
// With --gpus / --devices the frames are rendered by a pt_group (pt_launch_multi: all frames in one call, the image is
// bit-identical to the single-GPU loop); the same device may be listed more than once with the peer transport.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ptb200.h"

static void check(int rc)   // the reference's throw-and-catch-in-main convention (sutil/Exception.h, main :742)
{
    if (rc != PT_OK) throw std::runtime_error(pt_last_error());
}
static void cuda_check(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

int main(int argc, char** argv)
{
    std::string obj_path, out_path = "ptb_render.ppm", accum_out, accum_in;
    uint32_t width = 512, height = 512, spp = 128, depth = 4, frames = 1;   // :58-59, :43, :657
    bool direct = false, importance = false;                                // :655-656
    float bg[3] = {0.f, 0.f, 0.f};                                          // :568
    int pipeline = 0, transport = PT_TRANSPORT_PEER;
    std::vector<int> devices;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) throw std::runtime_error(std::string("missing value for ") + name);
            return argv[++i];
        };
        try {
            if (a == "--obj") obj_path = next("--obj");
            else if (a == "--width") width = (uint32_t)atoi(next("--width"));
            else if (a == "--height") height = (uint32_t)atoi(next("--height"));
            else if (a == "--spp") spp = (uint32_t)atoi(next("--spp"));
            else if (a == "--depth") depth = (uint32_t)atoi(next("--depth"));
            else if (a == "--frames") frames = (uint32_t)atoi(next("--frames"));
            else if (a == "--direct") direct = true;                       // key '0' (:111)
            else if (a == "--importance") importance = true;               // key '1' (:117)
            else if (a == "--out") out_path = next("--out");
            else if (a == "--accum-out") accum_out = next("--accum-out");
            else if (a == "--accum-in") accum_in = next("--accum-in");
            else if (a == "--pipeline") pipeline = atoi(next("--pipeline"));
            else if (a == "--gpus") {
                const int n = atoi(next("--gpus"));
                devices.clear();
                for (int d = 0; d < n; ++d) devices.push_back(d);
            } else if (a == "--devices") {
                devices.clear();
                std::string list = next("--devices");
                for (size_t pos = 0; pos <= list.size();) {
                    const size_t end = std::min(list.find(',', pos), list.size());
                    devices.push_back(atoi(list.substr(pos, end - pos).c_str()));
                    pos = end + 1;
                }
            } else if (a == "--transport") {
                const std::string t = next("--transport");
                if (t != "peer" && t != "nccl") throw std::runtime_error("--transport peer|nccl");
                transport = t == "nccl" ? PT_TRANSPORT_NCCL : PT_TRANSPORT_PEER;
            }
            else if (a == "--background") {
                for (int c = 0; c < 3; ++c) bg[c] = (float)atof(next("--background"));
            } else {
                fprintf(stderr, "unknown argument: %s\n", a.c_str());
                return 2;
            }
        } catch (const std::exception& e) {
            fprintf(stderr, "%s\n", e.what());
            return 2;
        }
    }
    if (obj_path.empty()) {
        fprintf(stderr, "usage: ptb_render --obj scene.obj [--width W --height H --spp N --depth D --direct --importance "
                        "--frames F --background r g b --out image.ppm --accum-out f --accum-in f --pipeline p]\n");
        return 2;
    }
    if (depth < 1) depth = 1;      // Up/Down keys clamp maxDepth to [1, 28] (:123-131)
    if (depth > 28) depth = 28;

    pt_obj* obj = nullptr;
    pt_scene* scene = nullptr;
    pt_group* group = nullptr;
    void *d_accum = nullptr, *d_frame = nullptr;
    int rc_main = 0;
    try {
        check(pt_obj_load(obj_path.c_str(), &obj));                                        // TinyObjWrapper obj(path) :650
        if (devices.empty()) {
            check(pt_scene_create_from_obj(obj, &scene));                                  // :671-683
        } else {
            // several GPUs: one replica of the scene per device, owned by the group
            check(pt_group_create_local(devices.data(), (uint32_t)devices.size(), transport, &group));
            uint32_t nv = 0, ni = 0, nm = 0, nmat = 0;
            check(pt_obj_counts(obj, &nv, &ni, &nm, &nmat));
            std::vector<float> verts(nv);
            std::vector<uint32_t> indices(ni), mat_ids(nm);
            std::vector<pt_material> mats(nmat);
            check(pt_obj_copy(obj, verts.data(), indices.data(), mat_ids.data(), mats.data()));
            check(pt_group_scene_create(group, verts.data(), nv / 4, indices.data(), ni / 3, mat_ids.data(), mats.data(), nmat, &scene));
            cuda_check(cudaSetDevice(devices[0]), "cudaSetDevice");
            printf("Group: %zu device(s), %s transport\n", devices.size(), transport == PT_TRANSPORT_NCCL ? "NCCL" : "peer-copy");
        }
        check(pt_scene_set_background(scene, bg[0], bg[1], bg[2]));
        check(pt_scene_set_pipeline(scene, pipeline));
        uint32_t n_tris = 0, n_nodes = 0;
        float build_ms = 0.f;
        check(pt_scene_info(scene, &n_tris, &n_nodes, &build_ms));
        printf("Scene: %u triangles, %u BVH nodes, build %.3f ms\n", n_tris, n_nodes, build_ms);

        pt_params params;
        memset(&params, 0, sizeof(params));
        params.width = width;
        params.height = height;
        params.samplesPerPixel = spp;
        params.maxDepth = depth;
        params.useDirectLighting = direct ? 1 : 0;
        params.useImportanceSampling = importance ? 1 : 0;
        const float eye[3] = {278.0f, 273.0f, -900.0f}, lookat[3] = {278.0f, 273.0f, 330.0f}, up[3] = {0.f, 1.f, 0.f};   // :228-233
        float U[3], V[3], W[3];
        pt_camera_uvw_frame(eye, lookat, up, 35.0f, (float)width / (float)height, U, V, W);                             // :664-666
        params.cameraEye = {eye[0], eye[1], eye[2]};
        params.cameraU = {U[0], U[1], U[2]};
        params.cameraV = {V[0], V[1], V[2]};
        params.cameraW = {W[0], W[1], W[2]};
        params.areaLight.emission = {10.0f, 10.0f, 10.0f};                                                                // :154-158
        params.areaLight.corner = {343.0f, 547.0f, 227.0f};
        params.areaLight.v1 = {0.0f, 0.0f, 105.0f};
        params.areaLight.v2 = {-130.0f, 0.0f, 0.0f};
        params.areaLight.normal = {0.0f, -1.0f, 0.0f};          // normalize(cross(v1, v2))
        params.handle = pt_scene_handle(scene);
        printf("Using Direct Lighting: %s\n", direct ? "yes" : "no");                                                     // :669-670
        printf("Using Importance Sampling: %s\n", importance ? "yes" : "no");

        const size_t px = (size_t)width * height;
        cuda_check(cudaMalloc(&d_accum, px * 16), "cudaMalloc(accumulation)");                                           // :145
        cuda_check(cudaMalloc(&d_frame, px * 4), "cudaMalloc(frame)");
        params.accumulationBuffer = d_accum;
        params.frameBuffer = d_frame;
        std::vector<float> h_accum(4 * px);
        std::vector<uint8_t> h_frame(4 * px);
        if (!accum_in.empty()) {                                 // resume: the running average + next frame index
            uint32_t next = 0;
            check(pt_accum_load(accum_in.c_str(), h_accum.data(), width, height, &next));
            cuda_check(cudaMemcpy(d_accum, h_accum.data(), px * 16, cudaMemcpyHostToDevice), "upload accumulation");
            params.currentFrameIdx = next;
            printf("Resuming at frame %u\n", next);
        }
        cudaStream_t stream;
        cuda_check(cudaStreamCreate(&stream), "cudaStreamCreate");                                                        // :161

        double total_ms = 0.0;
        unsigned long long sample_sum = 0;
        if (group && frames) {
            // the whole do-while of main (:700-730) in one collective call: frames dealt to the devices, slices folded in order
            const auto t0 = std::chrono::high_resolution_clock::now();
            check(pt_launch_multi(group, &params, frames, stream));
            cuda_check(cudaStreamSynchronize(stream), "render");
            params.currentFrameIdx += frames;
            total_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
            sample_sum = (unsigned long long)spp * frames;
            printf("Frame Render Time: %.3fms (average over %u frames on %zu devices)", total_ms / frames, frames, devices.size());
        }
        for (uint32_t f = 0; f < frames && !group; ++f) {                                                                // do ... while :700-730
            const auto t0 = std::chrono::high_resolution_clock::now();
            check(pt_launch(&params, stream));                                                                           // LaunchCurrentFrame :184
            cuda_check(cudaStreamSynchronize(stream), "render");                                                         // CUDA_SYNC_CHECK :209
            ++params.currentFrameIdx;                                                                                    // :714
            const auto t1 = std::chrono::high_resolution_clock::now();
            const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            total_ms += ms;
            sample_sum += spp;
            printf("\rFrame Render Time: %.3fms", ms);                                                                   // :726
            fflush(stdout);
        }
        printf("\nTotal Samples %llu\nAverage ms per frame: %.3f\nTotal ms: %.3f\n", sample_sum,                           // :738-740
               frames ? total_ms / frames : 0.0, total_ms);
        if (frames)
            printf("Throughput: %.1f Mpixel-samples/s\n", (double)px * (double)sample_sum / (total_ms * 1e-3) / 1e6);

        cuda_check(cudaMemcpy(h_frame.data(), d_frame, px * 4, cudaMemcpyDeviceToHost), "download frame");
        check(pt_image_write(out_path.c_str(), h_frame.data(), width, height));            // sutil::saveImage: .ppm or .png
        printf("Wrote %s\n", out_path.c_str());
        if (!accum_out.empty()) {
            cuda_check(cudaMemcpy(h_accum.data(), d_accum, px * 16, cudaMemcpyDeviceToHost), "download accumulation");
            check(pt_accum_save(accum_out.c_str(), h_accum.data(), width, height, params.currentFrameIdx));
            printf("Wrote %s (next frame %u)\n", accum_out.c_str(), params.currentFrameIdx);
        }
        cudaStreamDestroy(stream);
    } catch (const std::exception& e) {
        fprintf(stderr, "Caught exception: %s\n", e.what());                                                             // :742-744
        rc_main = 1;
    }
    if (d_accum) cudaFree(d_accum);                                                                                      // CleanAllTheThings :629
    if (d_frame) cudaFree(d_frame);
    if (group) pt_group_destroy(group);                                                                                 // destroys its scene replicas
    else if (scene) pt_scene_destroy(scene);
    if (obj) pt_obj_free(obj);
    return rc_main;
}
