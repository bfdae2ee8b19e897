// objload.cu -- OBJ/MTL ingest producing the four arrays the reference's TinyObjWrapper hands to the renderer
// (PathTracer_Optix/TinyObjWrapper.cpp:138-244): xyzw vertices, index triplets, per-triangle material ids and
// Material records.  Host code only.
//
// The reference delegates parsing to the vendored tinyobjloader v2.0.0 (util/tiny_obj_loader.h, MIT).  This is an
// independent single-pass reader for the subset that reaches those four arrays; it restates the published
// behaviour that determines the output bits:
//   * number syntax and rounding: digits accumulated in double, fractional digits weighted by 10^-k,
//     decimal exponent applied as ldexp(m * 5^e, e), then narrowed to float (tiny_obj_loader.h:897-1038);
//   * indices: 1-based, negative = relative to the vertices read so far, 0 is an error (:819-852, :1188-1238);
//   * triangles are emitted in file order; quads split along the shorter diagonal (:1511-1606); larger polygons
//     by ear clipping in the dominant plane (:1742-1960); faces with fewer than 3 corners are dropped;
//   * the material of a face is the last `usemtl` before it (-1 when none / unknown, :2604, :2849-2871);
//   * MTL: newmtl / Kd / Ke / Ni / Pr / Pm, defaults ior = 1, everything else 0 (:1386-1425, :2069-2300); a file
//     that defines no name still yields one unnamed material (:2458-2461);
//   * BSDF type from the material NAME (TinyObjWrapper.cpp:153-164).
// Shape/group names, normals, texture coordinates, lines and points do not reach the renderer and are skipped.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "pt_scene.h"

struct pt_obj {
    std::vector<float> verts;        // xyzw
    std::vector<uint32_t> indices;   // 3 per triangle
    std::vector<uint32_t> mat_ids;   // per triangle (0xffffffff = no material)
    std::vector<pt_material> mats;
    std::string warnings;
};

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// Decimal literal -> double with tinyobjloader's arithmetic (see header comment).  Returns false when the
// text is not a number; `out` is then left untouched.
bool parse_number(const char* s, const char* end, double* out)
{
    if (s >= end) return false;
    double m = 0.0;
    int expo = 0;
    char sign = '+', esign = '+';
    const char* c = s;
    int nread = 0;
    bool leading_dot = false;
    if (*c == '+' || *c == '-') {
        sign = *c++;
        if (c != end && *c == '.') leading_dot = true;
    } else if (is_digit(*c)) {
    } else if (*c == '.') {
        leading_dot = true;
    } else {
        return false;
    }
    bool more = c != end;
    if (!leading_dot) {
        while (more && is_digit(*c)) {
            m *= 10;
            m += (int)(*c - '0');
            ++c;
            ++nread;
            more = c != end;
        }
        if (nread == 0) return false;
    }
    if (more) {
        bool want_exp = false;
        if (*c == '.') {
            ++c;
            nread = 1;
            more = c != end;
            static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            while (more && is_digit(*c)) {
                m += (int)(*c - '0') * (nread < 8 ? lut[nread] : pow(10.0, -nread));
                ++nread;
                ++c;
                more = c != end;
            }
            want_exp = more;
        } else if (*c == 'e' || *c == 'E') {
            want_exp = true;
        }
        if (want_exp && (*c == 'e' || *c == 'E')) {
            ++c;
            more = c != end;
            if (more && (*c == '+' || *c == '-')) {
                esign = *c++;
            } else if (is_digit(*c)) {
            } else {
                return false;
            }
            nread = 0;
            more = c != end;
            while (more && is_digit(*c)) {
                if (expo > 2147483647 / 10) return false;
                expo *= 10;
                expo += (int)(*c - '0');
                ++c;
                ++nread;
                more = c != end;
            }
            expo *= (esign == '+' ? 1 : -1);
            if (nread == 0) return false;
        }
    }
    *out = (sign == '+' ? 1 : -1) * (expo ? ldexp(m * pow(5.0, expo), expo) : m);
    return true;
}

// next whitespace-delimited field as float; `dflt` when it is not a number
float next_float(const char*& p, double dflt = 0.0)
{
    p += strspn(p, " \t");
    const char* end = p + strcspn(p, " \t\r");
    double v = dflt;
    parse_number(p, end, &v);
    p = end;
    return (float)v;
}

std::string next_word(const char*& p)
{
    p += strspn(p, " \t");
    const size_t n = strcspn(p, " \t\r");
    std::string s(p, n);
    p += n;
    return s;
}

// Reads one line handling \n, \r\n and \r; false at end of input.
bool read_line(std::istream& in, std::string& line)
{
    line.clear();
    if (in.peek() == EOF) return false;
    std::streambuf* sb = in.rdbuf();
    for (;;) {
        const int c = sb->sbumpc();
        if (c == '\n') return true;
        if (c == '\r') {
            if (sb->sgetc() == '\n') sb->sbumpc();
            return true;
        }
        if (c == EOF) {
            if (line.empty()) in.setstate(std::ios::eofbit);
            return true;
        }
        line += (char)c;
    }
}

struct MtlRecord {
    std::string name;
    float Kd[3] = {0, 0, 0}, Ke[3] = {0, 0, 0};
    float Ni = 1.0f, Pr = 0.0f, Pm = 0.0f;
};

void read_mtl(std::istream& in, std::vector<MtlRecord>& mats, std::map<std::string, int>& by_name)
{
    MtlRecord cur;
    std::string line;
    while (read_line(in, line)) {
        const size_t last = line.find_last_not_of(" \t");
        line = last == std::string::npos ? std::string() : line.substr(0, last + 1);
        if (line.empty()) continue;
        const char* t = line.c_str();
        t += strspn(t, " \t");
        if (t[0] == '\0' || t[0] == '#') continue;
        if (strncmp(t, "newmtl", 6) == 0 && is_space(t[6])) {
            if (!cur.name.empty()) {
                by_name.insert(std::make_pair(cur.name, (int)mats.size()));
                mats.push_back(cur);
            }
            cur = MtlRecord();
            t += 7;
            cur.name = next_word(t);
            continue;
        }
        if (t[0] == 'K' && (t[1] == 'd' || t[1] == 'e') && is_space(t[2])) {
            float* dst = t[1] == 'd' ? cur.Kd : cur.Ke;
            t += 2;
            dst[0] = next_float(t);
            dst[1] = next_float(t);
            dst[2] = next_float(t);
            continue;
        }
        if (t[0] == 'N' && t[1] == 'i' && is_space(t[2])) {
            t += 2;
            cur.Ni = next_float(t);
            continue;
        }
        if (t[0] == 'P' && (t[1] == 'r' || t[1] == 'm') && is_space(t[2])) {
            float& dst = t[1] == 'r' ? cur.Pr : cur.Pm;
            t += 2;
            dst = next_float(t);
            continue;
        }
    }
    by_name.insert(std::make_pair(cur.name, (int)mats.size()));
    mats.push_back(cur);
}

struct Corner {
    int v;
};

// a/b/c, a//c, a/b, a -> vertex index (0-based); false on a zero or out-of-range relative index
bool parse_corner(const char*& p, int n_verts, int n_normals, int n_texcoords, Corner& out)
{
    auto fix = [](int idx, int n, bool allow_zero, int& ret) {
        if (idx > 0) {
            ret = idx - 1;
            return true;
        }
        if (idx == 0) {
            ret = -1;
            return allow_zero;
        }
        ret = n + idx;
        return ret >= 0;
    };
    int dummy;
    if (!fix(atoi(p), n_verts, false, out.v)) return false;
    p += strcspn(p, "/ \t\r");
    if (p[0] != '/') return true;
    ++p;
    if (p[0] == '/') {
        ++p;
        if (!fix(atoi(p), n_normals, true, dummy)) return false;
        p += strcspn(p, "/ \t\r");
        return true;
    }
    if (!fix(atoi(p), n_texcoords, true, dummy)) return false;
    p += strcspn(p, "/ \t\r");
    if (p[0] != '/') return true;
    ++p;
    if (!fix(atoi(p), n_normals, true, dummy)) return false;
    p += strcspn(p, "/ \t\r");
    return true;
}

// crossing-number point-in-triangle test used by the ear clipper
bool point_in_tri(const float* vx, const float* vy, float tx, float ty)
{
    bool c = false;
    for (int i = 0, j = 2; i < 3; j = i++) {
        if (((vy[i] > ty) != (vy[j] > ty)) && (tx < (vx[j] - vx[i]) * (ty - vy[i]) / (vy[j] - vy[i]) + vx[i])) c = !c;
    }
    return c;
}

struct Emitter {
    pt_obj* obj;
    const std::vector<float>* v;   // xyz per vertex read so far
    void tri(int a, int b, int c, int material)
    {
        obj->indices.push_back((uint32_t)a);
        obj->indices.push_back((uint32_t)b);
        obj->indices.push_back((uint32_t)c);
        obj->mat_ids.push_back((uint32_t)material);
    }
    bool in_range(int i) const { return (size_t)(3 * (size_t)(unsigned)i + 2) < v->size(); }

    void face(const std::vector<Corner>& f, int material)
    {
        const std::vector<float>& V = *v;
        size_t n = f.size();
        if (n < 3) {
            obj->warnings += "face with fewer than 3 corners dropped\n";
            return;
        }
        if (n == 3) {
            tri(f[0].v, f[1].v, f[2].v, material);
            return;
        }
        if (n == 4) {
            const int i0 = f[0].v, i1 = f[1].v, i2 = f[2].v, i3 = f[3].v;
            if (!in_range(i0) || !in_range(i1) || !in_range(i2) || !in_range(i3)) {
                obj->warnings += "quad with an out-of-range vertex index dropped\n";
                return;
            }
            const float* p0 = &V[3 * (size_t)i0];
            const float* p1 = &V[3 * (size_t)i1];
            const float* p2 = &V[3 * (size_t)i2];
            const float* p3 = &V[3 * (size_t)i3];
            const float ax = p2[0] - p0[0], ay = p2[1] - p0[1], az = p2[2] - p0[2];
            const float bx = p3[0] - p1[0], by = p3[1] - p1[1], bz = p3[2] - p1[2];
            const float d02 = ax * ax + ay * ay + az * az;
            const float d13 = bx * bx + by * by + bz * bz;
            if (d02 < d13) {
                tri(i0, i1, i2, material);
                tri(i0, i2, i3, material);
            } else {
                tri(i0, i1, i3, material);
                tri(i1, i2, i3, material);
            }
            return;
        }
        // ---- polygons with 5+ corners: ear clipping in the plane of the first non-degenerate corner ----
        size_t axes[2] = {1, 2};
        for (size_t k = 0; k < n; ++k) {
            const int a = f[k % n].v, b = f[(k + 1) % n].v, c = f[(k + 2) % n].v;
            if (!in_range(a) || !in_range(b) || !in_range(c)) continue;
            const float* pa = &V[3 * (size_t)a];
            const float* pb = &V[3 * (size_t)b];
            const float* pc = &V[3 * (size_t)c];
            const float e0x = pb[0] - pa[0], e0y = pb[1] - pa[1], e0z = pb[2] - pa[2];
            const float e1x = pc[0] - pb[0], e1y = pc[1] - pb[1], e1z = pc[2] - pb[2];
            const float cx = fabsf(e0y * e1z - e0z * e1y);
            const float cy = fabsf(e0z * e1x - e0x * e1z);
            const float cz = fabsf(e0x * e1y - e0y * e1x);
            const float eps = std::numeric_limits<float>::epsilon();
            if (cx > eps || cy > eps || cz > eps) {
                if (!(cx > cy && cx > cz)) {
                    axes[0] = 0;
                    if (cz > cx && cz > cy) axes[1] = 1;
                }
                break;
            }
        }
        std::vector<Corner> rest = f;
        size_t guess = 0;
        size_t budget = f.size();
        size_t prev_n = rest.size();
        while (rest.size() > 3 && budget > 0) {
            n = rest.size();
            if (guess >= n) guess -= n;
            if (prev_n != n) {
                prev_n = n;
                budget = n;
            } else {
                --budget;
            }
            Corner ind[3];
            float vx[3], vy[3];
            for (size_t k = 0; k < 3; ++k) {
                ind[k] = rest[(guess + k) % n];
                const size_t vi = (size_t)(unsigned)ind[k].v;
                if (vi * 3 + axes[0] >= V.size() || vi * 3 + axes[1] >= V.size()) {
                    vx[k] = vy[k] = 0.0f;
                } else {
                    vx[k] = V[vi * 3 + axes[0]];
                    vy[k] = V[vi * 3 + axes[1]];
                }
            }
            const float e0x = vx[1] - vx[0], e0y = vy[1] - vy[0];
            const float e1x = vx[2] - vx[1], e1y = vy[2] - vy[1];
            const float crs = e0x * e1y - e0y * e1x;
            const float area = (vx[0] * vy[1] - vy[0] * vx[1]) * 0.5f;
            if (crs * area < 0.0f) {   // reflex corner
                ++guess;
                continue;
            }
            bool overlap = false;
            for (size_t other = 3; other < n; ++other) {
                const size_t idx = (guess + other) % n;
                if (idx >= rest.size()) continue;
                const size_t ovi = (size_t)(unsigned)rest[idx].v;
                if (ovi * 3 + axes[0] >= V.size() || ovi * 3 + axes[1] >= V.size()) continue;
                if (point_in_tri(vx, vy, V[ovi * 3 + axes[0]], V[ovi * 3 + axes[1]])) {
                    overlap = true;
                    break;
                }
            }
            if (overlap) {
                ++guess;
                continue;
            }
            tri(ind[0].v, ind[1].v, ind[2].v, material);
            rest.erase(rest.begin() + (long)((guess + 1) % n));
        }
        if (rest.size() == 3) tri(rest[0].v, rest[1].v, rest[2].v, material);
    }
};

int bsdf_from_name(const std::string& name)   // TinyObjWrapper.cpp:153-164
{
    if (name.find("Refractive") != std::string::npos) return PT_BSDF_REFRACTION;
    if (name.find("Metallic") != std::string::npos) return PT_BSDF_METALLIC;
    return PT_BSDF_DIFFUSE;
}

}  // namespace

extern "C" {

int pt_obj_load(const char* path, pt_obj** out)
{
    if (!out) {
        ptb::set_error("pt_obj_load: out is NULL");
        return PT_ERR_INVALID;
    }
    *out = nullptr;
    if (!path) {
        ptb::set_error("pt_obj_load: path is NULL");
        return PT_ERR_INVALID;
    }
    std::ifstream in(path);
    if (!in) {
        ptb::set_error(std::string("Cannot open file [") + path + "]");
        return PT_ERR_IO;
    }
    std::string dir;
    {
        const std::string p(path);
        const size_t pos = p.find_last_of("/\\");
        if (pos != std::string::npos) dir = p.substr(0, pos);
    }
    pt_obj* obj = new pt_obj();
    std::vector<float> v;                 // xyz
    int n_normals = 0, n_texcoords = 0;
    std::vector<MtlRecord> mtl;
    std::map<std::string, int> mtl_by_name;
    std::set<std::string> mtl_files;
    int material = -1;
    // tinyobjloader buffers the faces of a group and triangulates them when the material changes, at g/o and at
    // the end of the file -- against ALL vertices read by then.  Quad splitting and ear clipping read coordinates,
    // so faces are buffered here the same way (a face may legally precede some `v` lines it does not use).
    std::vector<std::vector<Corner>> pending;
    Emitter emit{obj, &v};
    auto flush = [&]() {
        for (const auto& f : pending) emit.face(f, material);
        pending.clear();
    };
    std::string line;
    size_t line_no = 0;
    while (read_line(in, line)) {
        ++line_no;
        if (line.empty()) continue;
        const char* t = line.c_str();
        t += strspn(t, " \t");
        if (t[0] == '\0' || t[0] == '#') continue;
        if (t[0] == 'v' && is_space(t[1])) {
            t += 2;
            v.push_back(next_float(t));
            v.push_back(next_float(t));
            v.push_back(next_float(t));
            continue;
        }
        if (t[0] == 'v' && t[1] == 'n' && is_space(t[2])) {
            ++n_normals;
            continue;
        }
        if (t[0] == 'v' && t[1] == 't' && is_space(t[2])) {
            ++n_texcoords;
            continue;
        }
        if (t[0] == 'f' && is_space(t[1])) {
            t += 2;
            t += strspn(t, " \t");
            std::vector<Corner> f;
            while (!is_eol(t[0])) {
                Corner c;
                if (!parse_corner(t, (int)(v.size() / 3), n_normals, n_texcoords, c)) {
                    ptb::set_error("Failed to parse `f' line (e.g. a zero value for vertex index or invalid relative vertex index). Line " +
                                   std::to_string(line_no) + ").");
                    delete obj;
                    return PT_ERR_IO;
                }
                f.push_back(c);
                t += strspn(t, " \t\r");
            }
            pending.push_back(std::move(f));
            continue;
        }
        if (strncmp(t, "usemtl", 6) == 0) {
            t += 6;
            const std::string name = next_word(t);
            int id = -1;
            auto it = mtl_by_name.find(name);
            if (it != mtl_by_name.end()) id = it->second;
            else obj->warnings += "material [ '" + name + "' ] not found in .mtl\n";
            if (id != material) {
                flush();
                material = id;
            }
            continue;
        }
        if (strncmp(t, "mtllib", 6) == 0 && is_space(t[6])) {
            t += 7;
            // file names separated by blanks; backslash escapes a blank
            std::vector<std::string> names;
            {
                std::string tok;
                bool esc = false;
                for (const char* c = t; *c; ++c) {
                    if (esc) {
                        esc = false;
                    } else if (*c == '\\') {
                        esc = true;
                        continue;
                    } else if (*c == ' ') {
                        if (!tok.empty()) names.push_back(tok);
                        tok.clear();
                        continue;
                    }
                    tok += *c;
                }
                names.push_back(tok);
            }
            bool found = false;
            for (const std::string& nm : names) {
                if (mtl_files.count(nm)) {
                    found = true;
                    continue;
                }
                std::string full = nm;
                if (!dir.empty()) full = dir + (dir.back() == '/' ? "" : "/") + nm;
                std::ifstream min(full.c_str());
                if (min) {
                    read_mtl(min, mtl, mtl_by_name);
                    mtl_files.insert(nm);
                    found = true;
                    break;
                }
                obj->warnings += "Material file [ " + nm + " ] not found in a path : " + dir + "\n";
            }
            if (!found) obj->warnings += "Failed to load material file(s). Use default material.\n";
            continue;
        }
        if ((t[0] == 'g' || t[0] == 'o') && is_space(t[1])) {
            flush();
            continue;
        }
        // l, p, s, t, vw and unknown commands do not reach the renderer
    }
    flush();

    obj->verts.reserve(v.size() / 3 * 4);
    for (size_t i = 0; i + 2 < v.size(); i += 3) {   // TinyObjWrapper.cpp:200-221
        obj->verts.push_back(v[i]);
        obj->verts.push_back(v[i + 1]);
        obj->verts.push_back(v[i + 2]);
        obj->verts.push_back(1.0f);
    }
    obj->mats.resize(mtl.size());
    for (size_t m = 0; m < mtl.size(); ++m) {         // TinyObjWrapper.cpp:138-169
        pt_material& d = obj->mats[m];
        d.diffuse = pt_float3{mtl[m].Kd[0], mtl[m].Kd[1], mtl[m].Kd[2]};
        d.emission = pt_float3{mtl[m].Ke[0], mtl[m].Ke[1], mtl[m].Ke[2]};
        d.roughness = mtl[m].Pr;
        d.metallic = mtl[m].Pm;
        d.ior = mtl[m].Ni;
        d.bsdfType = bsdf_from_name(mtl[m].name);
    }
    *out = obj;
    return PT_OK;
}

void pt_obj_free(pt_obj* obj) { delete obj; }

int pt_obj_counts(const pt_obj* obj, uint32_t* n_vert_floats, uint32_t* n_indices, uint32_t* n_mat_ids, uint32_t* n_materials)
{
    if (!obj) {
        ptb::set_error("pt_obj_counts: obj is NULL");
        return PT_ERR_INVALID;
    }
    if (n_vert_floats) *n_vert_floats = (uint32_t)obj->verts.size();
    if (n_indices) *n_indices = (uint32_t)obj->indices.size();
    if (n_mat_ids) *n_mat_ids = (uint32_t)obj->mat_ids.size();
    if (n_materials) *n_materials = (uint32_t)obj->mats.size();
    return PT_OK;
}

int pt_obj_copy(const pt_obj* obj, float* verts_xyzw, uint32_t* indices, uint32_t* mat_ids, pt_material* materials)
{
    if (!obj) {
        ptb::set_error("pt_obj_copy: obj is NULL");
        return PT_ERR_INVALID;
    }
    if (verts_xyzw && !obj->verts.empty()) memcpy(verts_xyzw, obj->verts.data(), obj->verts.size() * sizeof(float));
    if (indices && !obj->indices.empty()) memcpy(indices, obj->indices.data(), obj->indices.size() * sizeof(uint32_t));
    if (mat_ids && !obj->mat_ids.empty()) memcpy(mat_ids, obj->mat_ids.data(), obj->mat_ids.size() * sizeof(uint32_t));
    if (materials && !obj->mats.empty()) memcpy(materials, obj->mats.data(), obj->mats.size() * sizeof(pt_material));
    return PT_OK;
}

int pt_scene_create_from_obj(const pt_obj* obj, pt_scene** out)
{
    if (!obj) {
        ptb::set_error("pt_scene_create_from_obj: obj is NULL");
        return PT_ERR_INVALID;
    }
    return pt_scene_create(obj->verts.data(), (uint32_t)(obj->verts.size() / 4), obj->indices.data(),
                           (uint32_t)(obj->indices.size() / 3), obj->mat_ids.data(), obj->mats.data(),
                           (uint32_t)obj->mats.size(), out);
}

}  // extern "C"
