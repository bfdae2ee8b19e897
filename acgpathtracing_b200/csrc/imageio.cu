// imageio.cu -- host-side image output and accumulation checkpoint (no device code).
//
//  * pt_image_write_ppm: what sutil::saveImage does for a UNSIGNED_BYTE4 buffer and a ".ppm" name
//    (sutil/sutil.cpp:543-570): binary P6, rows written top-down from a bottom-up framebuffer (row 0 of
//    frameBuffer is the bottom of the picture), alpha dropped.
//  * pt_accum_save / pt_accum_load: the running-average accumulation buffer plus the next currentFrameIdx is the
//    only resumable state of the renderer (pathTracerPrograms.cu:803-811); stored raw, little endian.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "pt_scene.h"

namespace {
const char kAccumMagic[8] = {'P', 'T', 'B', 'A', 'C', 'C', '1', 0};
int fail(int code, const std::string& msg)
{
    ptb::set_error(msg);
    return code;
}
}  // namespace

extern "C" {

int pt_image_write_ppm(const char* path, const void* frame_rgba8, uint32_t width, uint32_t height)
{
    if (!path || !frame_rgba8 || width == 0 || height == 0) return fail(PT_ERR_INVALID, "pt_image_write_ppm: bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(PT_ERR_IO, std::string("cannot open for writing: ") + path);
    fprintf(f, "P6\n%u %u\n255\n", width, height);
    const uint8_t* src = static_cast<const uint8_t*>(frame_rgba8);
    std::vector<uint8_t> row(3 * (size_t)width);
    bool ok = true;
    for (uint32_t j = height; j-- > 0 && ok;) {          // vertical flip
        const uint8_t* s = src + 4 * (size_t)width * j;
        for (uint32_t i = 0; i < width; ++i) {
            row[3 * i + 0] = s[4 * i + 0];
            row[3 * i + 1] = s[4 * i + 1];
            row[3 * i + 2] = s[4 * i + 2];
        }
        ok = fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? PT_OK : fail(PT_ERR_IO, std::string("short write: ") + path);
}

int pt_accum_save(const char* path, const void* accum_float4, uint32_t width, uint32_t height, uint32_t next_frame_idx)
{
    if (!path || !accum_float4 || width == 0 || height == 0) return fail(PT_ERR_INVALID, "pt_accum_save: bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(PT_ERR_IO, std::string("cannot open for writing: ") + path);
    const uint32_t hdr[3] = {width, height, next_frame_idx};
    const size_t bytes = 16 * (size_t)width * height;
    bool ok = fwrite(kAccumMagic, 1, 8, f) == 8 && fwrite(hdr, 4, 3, f) == 3 && fwrite(accum_float4, 1, bytes, f) == bytes;
    ok = (fclose(f) == 0) && ok;
    return ok ? PT_OK : fail(PT_ERR_IO, std::string("short write: ") + path);
}

int pt_accum_load(const char* path, void* accum_float4, uint32_t width, uint32_t height, uint32_t* next_frame_idx)
{
    if (!path || !accum_float4 || !next_frame_idx) return fail(PT_ERR_INVALID, "pt_accum_load: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(PT_ERR_IO, std::string("cannot open: ") + path);
    char magic[8];
    uint32_t hdr[3];
    const size_t bytes = 16 * (size_t)width * height;
    int rc = PT_OK;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, kAccumMagic, 8) != 0 || fread(hdr, 4, 3, f) != 3) {
        rc = fail(PT_ERR_IO, std::string("not an accumulation checkpoint: ") + path);
    } else if (hdr[0] != width || hdr[1] != height) {
        rc = fail(PT_ERR_INVALID, "checkpoint is " + std::to_string(hdr[0]) + "x" + std::to_string(hdr[1]) + ", expected " +
                                      std::to_string(width) + "x" + std::to_string(height));
    } else if (fread(accum_float4, 1, bytes, f) != bytes) {
        rc = fail(PT_ERR_IO, std::string("truncated checkpoint: ") + path);
    } else {
        *next_frame_idx = hdr[2];
    }
    fclose(f);
    return rc;
}

}  // extern "C"
