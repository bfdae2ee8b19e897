// imageio.cu -- host-side image output and accumulation checkpoint (no device code).
//
//  * pt_image_write_ppm: what sutil::saveImage does for a UNSIGNED_BYTE4 buffer and a ".ppm" name
//    (sutil/sutil.cpp:543-570): binary P6, rows written top-down from a bottom-up framebuffer (row 0 of
//    frameBuffer is the bottom of the picture), alpha dropped.
//  * pt_image_write_png: the ".png" branch of sutil::saveImage for the same buffer (sutil/sutil.cpp:622-638: stbi_write_png
//    with 4 components, flipped vertically): 8-bit RGBA, rows top-down.  The pixels are what matters, not the compression:
//    the zlib stream uses stored blocks (valid for every PNG reader, no deflate implementation to carry).
//  * pt_image_write: saveImage's dispatch on the last three characters of the file name (:548-549, :622).
//  * pt_accum_save / pt_accum_load: the running-average accumulation buffer plus the next currentFrameIdx is the
//    only resumable state of the renderer (pathTracerPrograms.cu:803-811); stored raw, little endian.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
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

int pt_image_write_png(const char* path, const void* frame_rgba8, uint32_t width, uint32_t height)
{
    if (!path || !frame_rgba8 || width == 0 || height == 0) return fail(PT_ERR_INVALID, "pt_image_write_png: bad argument");
    static uint32_t crc_table[256];
    static bool have_table = false;
    if (!have_table) {
        for (uint32_t n = 0; n < 256; ++n) {
            uint32_t c = n;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            crc_table[n] = c;
        }
        have_table = true;
    }
    auto be32 = [](std::vector<uint8_t>& v, uint32_t x) {
        for (int sft = 24; sft >= 0; sft -= 8) v.push_back((uint8_t)(x >> sft));
    };
    // raw image data: per row a filter byte 0 + RGBA, top row of the picture first (row height-1 of the frame buffer)
    const uint8_t* src = static_cast<const uint8_t*>(frame_rgba8);
    const size_t stride = 4 * (size_t)width + 1;
    std::vector<uint8_t> raw(stride * height);
    for (uint32_t j = 0; j < height; ++j) {
        raw[stride * j] = 0;
        memcpy(&raw[stride * j + 1], src + 4 * (size_t)width * (height - 1 - j), 4 * (size_t)width);
    }
    // zlib stream: header, stored deflate blocks of at most 65535 bytes, Adler-32 of the raw data
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78);
    z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xff));
        z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xff));
        z.push_back((uint8_t)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = pos; i < pos + n; ++i) {
            a = (a + raw[i]) % 65521u;
            b = (b + a) % 65521u;
        }
        pos += n;
    }
    be32(z, (b << 16) | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
        be32(out, (uint32_t)data.size());
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), data.begin(), data.end());
        uint32_t c = 0xffffffffu;
        for (size_t i = start; i < out.size(); ++i) c = crc_table[(c ^ out[i]) & 0xff] ^ (c >> 8);
        be32(out, c ^ 0xffffffffu);
    };
    std::vector<uint8_t> ihdr;
    be32(ihdr, width);
    be32(ihdr, height);
    const uint8_t tail[5] = {8, 6, 0, 0, 0};     // 8 bits per channel, colour type 6 (RGBA), deflate, no filter method, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    chunk("IHDR", ihdr);
    chunk("IDAT", z);
    chunk("IEND", {});
    FILE* f = fopen(path, "wb");
    if (!f) return fail(PT_ERR_IO, std::string("cannot open for writing: ") + path);
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (fclose(f) == 0) && ok;
    return ok ? PT_OK : fail(PT_ERR_IO, std::string("short write: ") + path);
}

int pt_image_write(const char* path, const void* frame_rgba8, uint32_t width, uint32_t height)
{
    const std::string name(path ? path : "");
    if (name.length() < 5) return fail(PT_ERR_INVALID, "pt_image_write: failed to determine filename extension");   // sutil.cpp:545-546
    const std::string ext = name.substr(name.length() - 3);
    if (ext == "PPM" || ext == "ppm") return pt_image_write_ppm(path, frame_rgba8, width, height);
    if (ext == "PNG" || ext == "png") return pt_image_write_png(path, frame_rgba8, width, height);
    return fail(PT_ERR_INVALID, "pt_image_write: unrecognized image file extension (ppm, png)");                     // sutil.cpp:... "Unrecognized image file extension"
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
