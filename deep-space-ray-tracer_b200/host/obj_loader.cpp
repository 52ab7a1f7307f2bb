// obj_loader.cpp -- Wavefront OBJ + MTL (+ map_Kd textures) -> the arrays dsrt_scene_upload takes.
//
// Host-side replacement for the reference's scene ingestion:
//   triangle_mesh::load_obj_from_file / load_mtl / to_cpu_material_from_mtl
//                                          /root/reference/inc/triangle_mesh.h:75-255
//   triangle::compute_normals              /root/reference/inc/triangle.h:70-73
//   upsert_material, collect_from_hittable, HostTextureRegistry, texture pool layout
//                                          /root/reference/src/gpu_scene_builder.cpp:72-138, 199-246,
//                                          258-282, 518-551
// The reference goes OBJ -> CPU scene graph (shared_ptr<material>, image_texture) -> flat
// GPUTriangle/GPUMaterial arrays; here the file is parsed straight into those arrays, with the
// same observable result (checked against the reference's own code in tests/test_obj_loader.py):
//   * `v` is parsed as double, scaled in double, then narrowed to float; `vt` is parsed as float
//     and stored as (u, 1 - v); `vn` is ignored -- every triangle gets the flat face normal
//     unit_vector(cross(v1 - v0, v2 - v0)) in float arithmetic, three times;
//   * faces are fan-triangulated around their first vertex; a face token is read with the
//     reference's sscanf patterns (v/vt/vn, v//vn, v/vt, v);
//   * a material is created the first time a face uses its name (later redefinitions of the name
//     do not change it) in the priority order emissive (Ke or map_Ke) > textured lambertian
//     (map_Kd) > dielectric (d < 0.999) > metal (|Ks| > 0.05, fuzz = 100/(Ns+100)) > lambertian
//     (Kd); the material table is ordered by first use, faces without a known material use the
//     reference main()'s fallback lambertian(0.73);
//   * a triangle whose current MTL entry has a map_Kd carries that texture's id and forces its
//     material's albedo to white; textures are registered by path in first-use order, converted
//     with powf(c / 255, 2.2) and packed into one float pool; an undecodable file becomes a 1x1
//     white texture;
//   * the reference's image_texture constructor switches stb_image's global "flip vertically on
//     load" flag on (inc/texture.h:133), and build_gpu_scene decodes the GPU textures after the
//     whole OBJ has been read -- so the pool holds the images bottom row first as soon as any
//     used material constructed an image_texture.  Reproduced.
// Differences, both where the reference has undefined behaviour: a face index outside
// [1, number of vertices read so far] (negative = relative indices included) is an unchecked
// vector access there; here the triangle is dropped and counted in `skipped_faces`.
#include "dsrt.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "image_decode.h"

namespace {

struct F3 { float x, y, z; };

struct MtlProps {                      // triangle_mesh.h:57-73
    F3 Kd{0.8f, 0.8f, 0.8f}, Ks{0.f, 0.f, 0.f}, Ke{0.f, 0.f, 0.f};
    double Ns = 0.0, d = 1.0, Ni = 1.5;
    std::string map_Kd, map_Ke;
};

// ---- std::istream-style token reading ----------------------------------------------------------
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

struct Cursor {
    const char* p;
    bool fail = false;
    explicit Cursor(const std::string& s) : p(s.c_str()) {}
    void skip() { while (*p && is_space(*p)) ++p; }
    // `iss >> std::string`: cleared first, fails when nothing is left
    bool word(std::string& out) {
        out.clear();
        if (fail) return false;
        skip();
        const char* b = p;
        while (*p && !is_space(*p)) ++p;
        if (p == b) { fail = true; return false; }
        out.assign(b, p);
        return true;
    }
    // the longest prefix that is a decimal floating literal: [sign] digits [. digits] [e [sign] digits]
    bool number_span(const char*& b, const char*& e) {
        if (fail) return false;
        skip();
        b = p;
        const char* q = p;
        if (*q == '+' || *q == '-') ++q;
        const char* d0 = q;
        while (*q >= '0' && *q <= '9') ++q;
        size_t digits = (size_t)(q - d0);
        if (*q == '.') { ++q; const char* d1 = q; while (*q >= '0' && *q <= '9') ++q; digits += (size_t)(q - d1); }
        if (digits == 0) { fail = true; return false; }
        if (*q == 'e' || *q == 'E') {
            const char* x = q + 1;
            if (*x == '+' || *x == '-') ++x;
            if (*x >= '0' && *x <= '9') { while (*x >= '0' && *x <= '9') ++x; q = x; }
        }
        e = q;
        p = q;
        return true;
    }
    bool real(double& v) {                 // `iss >> double`
        const char *b, *e;
        if (!number_span(b, e)) return false;
        v = std::strtod(std::string(b, e).c_str(), nullptr);
        return true;
    }
    bool realf(float& v) {                 // `iss >> float` (rounded once, decimal -> float)
        const char *b, *e;
        if (!number_span(b, e)) return false;
        v = std::strtof(std::string(b, e).c_str(), nullptr);
        return true;
    }
};

// parse_face_idx, triangle_mesh.h:181-188
void parse_face_idx(const std::string& t, int& v, int& vt, int& vn) {
    v = vt = vn = 0;
    int a = 0, b = 0, c = 0;
    if (std::sscanf(t.c_str(), "%d/%d/%d", &a, &b, &c) == 3) { v = a; vt = b; vn = c; return; }
    a = b = c = 0;
    if (std::sscanf(t.c_str(), "%d//%d", &a, &c) == 2) { v = a; vn = c; return; }
    a = b = c = 0;
    if (std::sscanf(t.c_str(), "%d/%d", &a, &b) == 2) { v = a; vt = b; return; }
    a = 0;
    if (std::sscanf(t.c_str(), "%d", &a) == 1) { v = a; return; }
}

// load_mtl, triangle_mesh.h:119-170
void load_mtl(const std::string& path, std::unordered_map<std::string, MtlProps>& table) {
    std::ifstream in(path);
    if (!in) return;
    std::unordered_map<std::string, MtlProps> out;
    std::string line, tag, cur;
    MtlProps props;
    auto flush = [&]() { if (!cur.empty()) out[cur] = props; };
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        Cursor c(line);
        if (!c.word(tag)) continue;
        double r, g, b, s;
        if (tag == "newmtl") { flush(); props = MtlProps{}; c.word(cur); }
        else if (tag == "Kd") { if (c.real(r) && c.real(g) && c.real(b)) props.Kd = F3{(float)r, (float)g, (float)b}; }
        else if (tag == "Ks") { if (c.real(r) && c.real(g) && c.real(b)) props.Ks = F3{(float)r, (float)g, (float)b}; }
        else if (tag == "Ke") { if (c.real(r) && c.real(g) && c.real(b)) props.Ke = F3{(float)r, (float)g, (float)b}; }
        else if (tag == "Ns") { if (c.real(s)) props.Ns = s; }
        else if (tag == "d") { if (c.real(s)) props.d = s; }
        else if (tag == "Ni") { if (c.real(s)) props.Ni = s; }
        else if (tag == "map_Kd") { c.word(props.map_Kd); }
        else if (tag == "map_Ke") { c.word(props.map_Ke); }
    }
    flush();
    for (auto& kv : out) table[kv.first] = kv.second;
}

// to_cpu_material_from_mtl (triangle_mesh.h:75-117) followed by upsert_material
// (gpu_scene_builder.cpp:72-138).  `made_image_texture` reports whether the reference would
// have constructed an image_texture (which sets stb's vertical-flip flag).
DsrtMaterial material_from_mtl(const MtlProps& m, bool& made_image_texture) {
    DsrtMaterial g;
    std::memset(&g, 0, sizeof(g));
    g.albedo_tex = -1;
    made_image_texture = false;
    const bool has_emissive = (m.Ke.x != 0.0f || m.Ke.y != 0.0f || m.Ke.z != 0.0f);
    if (has_emissive || !m.map_Ke.empty()) {
        g.type = DSRT_MAT_DIFFUSE_LIGHT;
        g.albedo = DsrtVec3{1.f, 1.f, 1.f};
        if (!m.map_Ke.empty()) { made_image_texture = true; g.emissive = DsrtVec3{1.f, 1.f, 1.f}; }   // emit_value() of a textured light
        else g.emissive = DsrtVec3{m.Ke.x, m.Ke.y, m.Ke.z};
        g.fuzz = 0.f; g.ref_idx = 1.f;
        return g;
    }
    if (!m.map_Kd.empty()) {
        made_image_texture = true;
        g.type = DSRT_MAT_LAMBERTIAN;
        g.albedo = DsrtVec3{0.8f, 0.8f, 0.8f};            // lambertian::albedo_value() of a textured albedo
        g.fuzz = 0.f; g.ref_idx = 1.5f;
        return g;
    }
    if (m.d < 0.999) {
        const double ior = (m.Ni > 0.1 && m.Ni < 10.0) ? m.Ni : 1.5;
        g.type = DSRT_MAT_DIELECTRIC;
        g.albedo = DsrtVec3{1.f, 1.f, 1.f};
        g.fuzz = 0.f; g.ref_idx = (float)ior;
        return g;
    }
    const double ks_mag = (double)sqrtf(m.Ks.x * m.Ks.x + m.Ks.y * m.Ks.y + m.Ks.z * m.Ks.z);
    if (ks_mag > 0.05) {
        double fuzz = 100.0 / (m.Ns + 100.0);
        fuzz = fuzz < 0.0 ? 0.0 : (fuzz > 1.0 ? 1.0 : fuzz);      // std::clamp; NaN passes through as in the reference
        fuzz = fuzz < 1 ? fuzz : 1;                                // metal::metal
        g.type = DSRT_MAT_METAL;
        g.albedo = DsrtVec3{m.Ks.x, m.Ks.y, m.Ks.z};
        g.fuzz = (float)fuzz; g.ref_idx = 1.5f;
        return g;
    }
    g.type = DSRT_MAT_LAMBERTIAN;
    g.albedo = DsrtVec3{m.Kd.x, m.Kd.y, m.Kd.z};
    g.fuzz = 0.f; g.ref_idx = 1.5f;
    return g;
}

// unit_vector(cross(v1 - v0, v2 - v0)) with the reference's float expressions (inc/vec3.h)
DsrtVec3 face_normal(F3 v0, F3 v1, F3 v2) {
    const float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z;
    const float bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
    const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    const float len = sqrtf(cx * cx + cy * cy + cz * cz);
    const float inv = 1.0f / len;
    return DsrtVec3{cx * inv, cy * inv, cz * inv};
}

}  // namespace

struct dsrt_mesh {
    std::vector<DsrtTriangle> tris;
    std::vector<DsrtMaterial> mats;
    std::vector<DsrtTextureHeader> tex;
    std::vector<float> pool;
    std::vector<std::string> tex_paths;
    int nverts = 0, nuvs = 0, skipped = 0, tex_failed = 0, tex_unsupported = 0, flipped = 0;
};

static int mesh_load_obj_impl(const char* obj_path, double scale, dsrt_mesh** out, std::unique_ptr<dsrt_mesh>& holder) {
    std::ifstream in(obj_path);
    if (!in) return DSRT_EIO;
    holder.reset(new dsrt_mesh());
    dsrt_mesh* M = holder.get();

    const std::string path(obj_path);
    const size_t slash = path.find_last_of("/\\");
    const std::string base_dir = path.substr(0, slash == std::string::npos ? 0 : slash + 1);

    std::vector<F3> verts, uvs;
    std::unordered_map<std::string, MtlProps> mtl;
    std::unordered_map<std::string, int> mat_cache;        // material name -> table index, fixed at first use
    std::unordered_map<std::string, int> tex_index;        // texture path -> id
    std::vector<int> tri_tex_ref;                          // per triangle: index into M->tex_paths or -1
    int fallback_index = -1;
    bool flip = false;

    std::string line, tag, cur_mtl;
    std::vector<std::string> toks;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        Cursor c(line);
        if (!c.word(tag)) continue;
        if (tag == "mtllib") {
            std::string name; c.word(name);
            load_mtl(base_dir + name, mtl);
        } else if (tag == "usemtl") {
            c.word(cur_mtl);
        } else if (tag == "v") {
            double x, y, z;
            if (c.real(x) && c.real(y) && c.real(z)) verts.push_back(F3{(float)(scale * x), (float)(scale * y), (float)(scale * z)});
        } else if (tag == "vt") {
            float u, v;
            if (c.realf(u) && c.realf(v)) uvs.push_back(F3{u, 1.0f - v, 0.0f});
        } else if (tag == "f") {
            toks.clear();
            std::string tok;
            while (c.word(tok)) toks.push_back(tok);
            if (toks.size() < 3) continue;

            int mat = -1;
            if (!cur_mtl.empty()) {
                auto it = mat_cache.find(cur_mtl);
                if (it != mat_cache.end()) mat = it->second;
                else {
                    auto mit = mtl.find(cur_mtl);
                    if (mit != mtl.end()) {
                        bool img = false;
                        M->mats.push_back(material_from_mtl(mit->second, img));
                        if (img) flip = true;
                        mat = (int)M->mats.size() - 1;
                        mat_cache[cur_mtl] = mat;
                    }
                }
            }
            // Note: the reference assigns GPU material ids when it later walks the triangles
            // (first use by a triangle), which is the order materials are created here as long
            // as every face that creates one also emits a triangle; `order_fixup` below restores
            // that order when a face created a material but produced no triangle.
            int i0, it0, in0;
            parse_face_idx(toks[0], i0, it0, in0);
            if (i0 == 0) continue;
            for (size_t k = 1; k + 1 < toks.size(); ++k) {
                int i1, it1, in1, i2, it2, in2;
                parse_face_idx(toks[k], i1, it1, in1);
                parse_face_idx(toks[k + 1], i2, it2, in2);
                if (i1 == 0 || i2 == 0) continue;
                const int nv = (int)verts.size();
                if (i0 < 1 || i1 < 1 || i2 < 1 || i0 > nv || i1 > nv || i2 > nv) { ++M->skipped; continue; }
                if (mat < 0) {
                    if (fallback_index < 0) {
                        DsrtMaterial g;
                        std::memset(&g, 0, sizeof(g));
                        g.type = DSRT_MAT_LAMBERTIAN; g.albedo_tex = -1;
                        g.albedo = DsrtVec3{0.73f, 0.73f, 0.73f};          // main.cpp:241
                        g.ref_idx = 1.5f;
                        M->mats.push_back(g);
                        fallback_index = (int)M->mats.size() - 1;
                    }
                    mat = -2;      // resolved to fallback_index below; not cached under the name
                }
                const F3 v0 = verts[i0 - 1], v1 = verts[i1 - 1], v2 = verts[i2 - 1];
                const F3 z{0.f, 0.f, 0.f};
                const F3 uv0 = (it0 > 0 && it0 <= (int)uvs.size()) ? uvs[it0 - 1] : z;
                const F3 uv1 = (it1 > 0 && it1 <= (int)uvs.size()) ? uvs[it1 - 1] : z;
                const F3 uv2 = (it2 > 0 && it2 <= (int)uvs.size()) ? uvs[it2 - 1] : z;
                DsrtTriangle t;
                std::memset(&t, 0, sizeof(t));
                t.v0 = DsrtVec3{v0.x, v0.y, v0.z}; t.v1 = DsrtVec3{v1.x, v1.y, v1.z}; t.v2 = DsrtVec3{v2.x, v2.y, v2.z};
                t.n0 = t.n1 = t.n2 = face_normal(v0, v1, v2);
                t.uv0 = DsrtVec3{uv0.x, uv0.y, 0.f}; t.uv1 = DsrtVec3{uv1.x, uv1.y, 0.f}; t.uv2 = DsrtVec3{uv2.x, uv2.y, 0.f};
                t.material_id = mat == -2 ? fallback_index : mat;
                t.albedo_tex = -1;
                int tref = -1;
                if (!cur_mtl.empty()) {
                    auto mit = mtl.find(cur_mtl);
                    if (mit != mtl.end() && !mit->second.map_Kd.empty()) {
                        const std::string tp = base_dir + mit->second.map_Kd;
                        auto ti = tex_index.find(tp);
                        if (ti == tex_index.end()) {
                            tref = (int)M->tex_paths.size();
                            tex_index[tp] = tref;
                            M->tex_paths.push_back(tp);
                        } else tref = ti->second;
                    }
                }
                tri_tex_ref.push_back(tref);
                M->tris.push_back(t);
            }
        }
    }
    M->nverts = (int)verts.size();
    M->nuvs = (int)uvs.size();

    // Material ids in the reference follow first use by a TRIANGLE (upsert_material is called
    // while walking mesh->triangles).  Materials created above by faces that emitted no triangle
    // would be out of that order (or absent): renumber by first use and drop the unused.
    {
        std::vector<int> remap(M->mats.size(), -1);
        std::vector<DsrtMaterial> ordered;
        for (DsrtTriangle& t : M->tris) {
            if (remap[(size_t)t.material_id] < 0) {
                remap[(size_t)t.material_id] = (int)ordered.size();
                ordered.push_back(M->mats[(size_t)t.material_id]);
            }
            t.material_id = remap[(size_t)t.material_id];
        }
        M->mats.swap(ordered);
    }

    // Textures: decoded after the whole OBJ has been read (as build_gpu_scene does), so the flip
    // flag has its final value for every texture.
    M->flipped = flip ? 1 : 0;
    int float_offset = 0;
    for (const std::string& tp : M->tex_paths) {
        dsrt_host::DecodedImage img;
        DsrtTextureHeader h;
        bool exists = false;
        if (!dsrt_host::decode_image_file(tp.c_str(), img, &exists)) {
            std::fprintf(stderr, "WARN: failed to load texture '%s'%s\n", tp.c_str(), exists ? " (the file exists but is in a format this loader does not decode)" : "");
            ++M->tex_failed;
            if (exists) ++M->tex_unsupported;
            h.width = 1; h.height = 1; h.offset = float_offset;
            M->pool.insert(M->pool.end(), {1.0f, 1.0f, 1.0f});
            float_offset += 3;
        } else {
            h.width = img.w; h.height = img.h; h.offset = float_offset;
            const size_t row = (size_t)img.w * 3;
            M->pool.reserve(M->pool.size() + row * (size_t)img.h);
            for (int y = 0; y < img.h; ++y) {
                const unsigned char* src = &img.rgb[row * (size_t)(flip ? img.h - 1 - y : y)];
                for (size_t k = 0; k < row; ++k) M->pool.push_back(powf((float)src[k] / 255.0f, 2.2f));   // sRGB -> linear, :224-226
            }
            float_offset += (int)(row * (size_t)img.h);
        }
        M->tex.push_back(h);
    }
    for (size_t i = 0; i < M->tris.size(); ++i) {
        if (tri_tex_ref[i] < 0) continue;
        M->tris[i].albedo_tex = tri_tex_ref[i];
        M->mats[(size_t)M->tris[i].material_id].albedo = DsrtVec3{1.0f, 1.0f, 1.0f};      // :275-278
    }
    *out = holder.release();
    return DSRT_OK;
}

// No exception crosses the C ABI: a malformed or huge file ends as an error code.
extern "C" int dsrt_mesh_load_obj(const char* obj_path, double scale, dsrt_mesh** out) {
    if (!obj_path || !out) return DSRT_EINVAL;
    *out = nullptr;
    std::unique_ptr<dsrt_mesh> holder;
    try {
        return mesh_load_obj_impl(obj_path, scale, out, holder);
    } catch (const std::bad_alloc&) {
        return DSRT_ENOMEM;
    } catch (...) {
        return DSRT_EIO;
    }
}

extern "C" int dsrt_mesh_view(const dsrt_mesh* m, DsrtMeshView* v) {
    if (!m || !v) return DSRT_EINVAL;
    v->triangles = m->tris.data(); v->num_triangles = (int)m->tris.size();
    v->materials = m->mats.data(); v->num_materials = (int)m->mats.size();
    v->textures = m->tex.data(); v->num_textures = (int)m->tex.size();
    v->texture_pool = m->pool.data(); v->texture_pool_floats = (int)m->pool.size();
    v->num_vertices = m->nverts; v->num_uvs = m->nuvs;
    v->skipped_faces = m->skipped; v->textures_failed = m->tex_failed; v->textures_flipped = m->flipped;
    v->textures_unsupported = m->tex_unsupported;
    return DSRT_OK;
}

extern "C" void dsrt_mesh_free(dsrt_mesh* m) { delete m; }
