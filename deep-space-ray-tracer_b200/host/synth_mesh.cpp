// synth_mesh.cpp -- deterministic synthetic "station" mesh standing in for the ISS OBJ,
// which is not part of the reference tree.
//
// The station is assembled from parametric patches (truss boxes, module cylinders with end
// caps, solar-array and radiator sheets, small equipment boxes and dishes) inside
// x in [-950,950], y in [-420,420], z in [-25,25] model units, tessellated so that exactly
// `ntris` triangles come out.  Patch densities are drawn log-uniformly so triangle sizes
// span more than a decade, like a CAD export.  Every vertex gets a tiny deterministic
// displacement and the whole station a sub-degree rotation, so no triangle is axis aligned:
// the reference's slab test rejects boxes of zero thickness (gpu_render.cu:309-312,
// `t_max <= t_min`), i.e. axis-aligned coplanar leaves are invisible to it, and parity
// scenes must stay clear of that quirk ("reference quirks").
//
// Face normals are computed with the reference's float expressions
// (inc/triangle.h:70-73, inc/vec3.h:101-129) so that the reference's own `triangle`
// constructor reproduces them bit for bit from the same vertices.
#include "dsrt.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct D3 { double x, y, z; };
inline D3 d3(double x, double y, double z) { return D3{x, y, z}; }
inline D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
inline D3 crossd(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline double lend(D3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline D3 normd(D3 a) { double l = lend(a); return l > 0 ? a * (1.0 / l) : a; }

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull) {}
    uint64_t next() {   // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double range(double a, double b) { return a + (b - a) * uni(); }
};

inline double hash01(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    Rng r(a * 0x100000001B3ull ^ (b + 0x9E37ull) * 0xC2B2AE3D27D4EB4Full ^ (c + 77) * 0x165667B19E3779F9ull ^ d);
    r.next();
    return r.uni();
}

enum PatchKind { QUAD, CYL, DISC };

struct Patch {
    PatchKind kind;
    D3 o, a, b;        // QUAD: origin and the two edge vectors
                       // CYL : o = centre, a = unit axis, b = (radius, length, 0)
                       // DISC: o = centre, a = unit normal, b = (radius, inner fraction, 0)
    double area;
    double density;    // relative triangles per unit area
    int mat;
    bool wrap_u;
};

void basis(D3 axis, D3& e1, D3& e2) {
    D3 h = std::fabs(axis.x) > 0.9 ? d3(0, 1, 0) : d3(1, 0, 0);
    e1 = normd(crossd(axis, h));
    e2 = crossd(axis, e1);
}

D3 eval(const Patch& p, double u, double v) {
    switch (p.kind) {
    case QUAD: return p.o + p.a * u + p.b * v;
    case CYL: {
        D3 e1, e2; basis(p.a, e1, e2);
        double ang = 2.0 * M_PI * u;
        return p.o + p.a * ((v - 0.5) * p.b.y) + (e1 * std::cos(ang) + e2 * std::sin(ang)) * p.b.x;
    }
    default: {
        D3 e1, e2; basis(p.a, e1, e2);
        double ang = 2.0 * M_PI * u;
        double r = p.b.x * (p.b.y + (1.0 - p.b.y) * v);
        return p.o + (e1 * std::cos(ang) + e2 * std::sin(ang)) * r;
    }
    }
}

void patch_extent(const Patch& p, double& lu, double& lv) {
    switch (p.kind) {
    case QUAD: lu = lend(p.a); lv = lend(p.b); break;
    case CYL:  lu = 2.0 * M_PI * p.b.x; lv = p.b.y; break;
    default:   lu = 2.0 * M_PI * p.b.x * 0.5 * (1.0 + p.b.y); lv = p.b.x * (1.0 - p.b.y); break;
    }
}

struct Builder {
    std::vector<Patch> patches;
    Rng rng;
    explicit Builder(uint32_t seed) : rng(seed) {}

    double log_density() { return std::exp2(rng.range(-2.0, 2.0)); }

    void quad(D3 o, D3 a, D3 b, int mat, double dens_scale = 1.0) {
        Patch p{QUAD, o, a, b, lend(crossd(a, b)), log_density() * dens_scale, mat, false};
        patches.push_back(p);
    }
    void box(D3 c, D3 ex, D3 ey, D3 ez, int mat, double dens_scale = 1.0) {
        // ex, ey, ez are half-extent vectors (need not be axis aligned)
        double dens = log_density() * dens_scale;
        const D3 h[3] = {ex, ey, ez};
        for (int ax = 0; ax < 3; ++ax)
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                D3 n = h[ax] * (double)sgn, a = h[(ax + 1) % 3], b = h[(ax + 2) % 3];
                Patch p{QUAD, c + n - a - b, a * 2.0, b * 2.0, 4.0 * lend(crossd(a, b)), dens, mat, false};
                patches.push_back(p);
            }
    }
    void cylinder(D3 c, D3 axis, double radius, double length, int mat, bool caps, double dens_scale = 1.0) {
        axis = normd(axis);
        double dens = log_density() * dens_scale;
        Patch p{CYL, c, axis, d3(radius, length, 0), 2.0 * M_PI * radius * length, dens, mat, true};
        patches.push_back(p);
        if (caps)
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                Patch q{DISC, c + axis * (0.5 * length * sgn), axis, d3(radius, 0.05, 0),
                        M_PI * radius * radius, dens, mat, true};
                patches.push_back(q);
            }
    }
};

// rotation helpers (degrees)
D3 rot_x(D3 p, double deg) { double r = deg * M_PI / 180.0, c = std::cos(r), s = std::sin(r); return d3(p.x, c * p.y - s * p.z, s * p.y + c * p.z); }
D3 rot_y(D3 p, double deg) { double r = deg * M_PI / 180.0, c = std::cos(r), s = std::sin(r); return d3(c * p.x + s * p.z, p.y, -s * p.x + c * p.z); }
D3 rot_z(D3 p, double deg) { double r = deg * M_PI / 180.0, c = std::cos(r), s = std::sin(r); return d3(c * p.x - s * p.y, s * p.x + c * p.y, p.z); }

// Material table (indices used below).
enum {
    M_TRUSS = 0, M_WHITE, M_DARK, M_GOLD, M_BLUE, M_RED, M_GREEN,       // lambertian
    M_STEEL, M_FOIL, M_BRUSHED,                                         // metal
    M_GLASS,                                                            // dielectric
    M_COUNT
};

void fill_materials(DsrtMaterial* mats, int diffuse_only) {
    struct Def { int type; float r, g, b, fuzz, ior; };
    const Def defs[M_COUNT] = {
        {DSRT_MAT_LAMBERTIAN, 0.73f, 0.73f, 0.73f, 0.f, 1.5f},
        {DSRT_MAT_LAMBERTIAN, 0.80f, 0.80f, 0.78f, 0.f, 1.5f},
        {DSRT_MAT_LAMBERTIAN, 0.20f, 0.20f, 0.22f, 0.f, 1.5f},
        {DSRT_MAT_LAMBERTIAN, 0.75f, 0.60f, 0.25f, 0.f, 1.5f},
        {DSRT_MAT_LAMBERTIAN, 0.15f, 0.20f, 0.55f, 0.f, 1.5f},
        {DSRT_MAT_LAMBERTIAN, 0.65f, 0.25f, 0.20f, 0.f, 1.5f},
        {DSRT_MAT_LAMBERTIAN, 0.30f, 0.55f, 0.30f, 0.f, 1.5f},
        {DSRT_MAT_METAL,      0.80f, 0.80f, 0.85f, 0.10f, 1.5f},
        {DSRT_MAT_METAL,      0.85f, 0.70f, 0.35f, 0.30f, 1.5f},
        {DSRT_MAT_METAL,      0.60f, 0.60f, 0.60f, 0.50f, 1.5f},
        {DSRT_MAT_DIELECTRIC, 1.00f, 1.00f, 1.00f, 0.f, 1.5f},
    };
    for (int i = 0; i < M_COUNT; ++i) {
        DsrtMaterial m;
        std::memset(&m, 0, sizeof(m));
        m.type = diffuse_only ? DSRT_MAT_LAMBERTIAN : defs[i].type;
        m.albedo_tex = -1;
        // diffuse-only keeps each part's colour but glass becomes light grey
        m.albedo.x = defs[i].r; m.albedo.y = defs[i].g; m.albedo.z = defs[i].b;
        if (diffuse_only && defs[i].type == DSRT_MAT_DIELECTRIC) { m.albedo.x = m.albedo.y = m.albedo.z = 0.6f; }
        m.fuzz = diffuse_only ? 0.f : defs[i].fuzz;
        m.ref_idx = (!diffuse_only && defs[i].type == DSRT_MAT_DIELECTRIC) ? defs[i].ior : 1.5f;
        mats[i] = m;
    }
}

void build_station(Builder& B) {
    Rng& r = B.rng;
    // --- main truss along x: 18 box segments, slightly different cross-sections
    for (int i = 0; i < 18; ++i) {
        double x0 = -900.0 + 100.0 * i;
        double hy = r.range(9.0, 12.0), hz = r.range(8.0, 11.0);
        int mat = (i % 5 == 0) ? M_STEEL : (i % 3 == 0 ? M_WHITE : M_TRUSS);
        B.box(d3(x0 + 50.0, r.range(-1.5, 1.5), r.range(-1.0, 1.0)), d3(49.0, 0, 0), d3(0, hy, 0), d3(0, 0, hz), mat);
    }
    // --- pressurised modules: a spine along y through the centre and a cross-piece along x
    for (int i = 0; i < 8; ++i) {
        double y0 = -300.0 + 75.0 * i;
        double rad = r.range(13.0, 18.0);
        int mat = (i % 3 == 0) ? M_FOIL : (i % 3 == 1 ? M_WHITE : M_BRUSHED);
        B.cylinder(d3(r.range(-2, 2), y0 + 37.5, r.range(-2, 2)), d3(0.01 * r.range(-1, 1), 1, 0.01 * r.range(-1, 1)), rad, 72.0, mat, true, 1.5);
    }
    for (int i = 0; i < 6; ++i) {
        double x0 = -240.0 + 80.0 * i;
        double rad = r.range(11.0, 16.0);
        int mat = (i % 2 == 0) ? M_WHITE : M_STEEL;
        B.cylinder(d3(x0 + 40.0, -70.0 + r.range(-2, 2), r.range(-3, 3)), d3(1, 0.01 * r.range(-1, 1), 0.01 * r.range(-1, 1)), rad, 76.0, mat, true, 1.5);
    }
    // --- solar array wings: long sheets along +-y, feathered about their long axis
    const double wing_x[4] = {520.0, 660.0, 790.0, 905.0};
    for (int side = -1; side <= 1; side += 2)
        for (int k = 0; k < 4; ++k)
            for (int ysgn = -1; ysgn <= 1; ysgn += 2) {
                double alpha = r.range(-38.0, 38.0);
                double halfw = r.range(28.0, 34.0);
                D3 c = d3(side * wing_x[k], ysgn * 225.0, 0);
                D3 a = rot_y(d3(2.0 * halfw, 0, 0), alpha);          // width, feathered about y
                D3 b = d3(0, 370.0, 0);                               // length along y
                // two blankets with a gap for the mast
                for (int blanket = 0; blanket < 2; ++blanket) {
                    D3 off = a * (blanket == 0 ? -0.52 : 0.02);
                    B.quad(c + off - b * 0.5, a * 0.5, b, (blanket + k) % 4 == 0 ? M_GLASS : M_BLUE, 2.0);
                }
                // mast
                B.cylinder(c, d3(0, 1, 0), 1.2, 370.0, M_STEEL, false, 0.6);
            }
    // --- radiators: sheets along x near the centre, feathered about x
    for (int i = 0; i < 6; ++i) {
        double alpha = r.range(25.0, 55.0) * ((i & 1) ? -1.0 : 1.0);
        double cx = (i < 3 ? -1.0 : 1.0) * (150.0 + 110.0 * (i % 3));
        double cy = ((i & 1) ? -1.0 : 1.0) * r.range(95.0, 170.0);
        D3 a = d3(95.0, 0, 0);
        D3 b = rot_x(d3(0, 38.0, 0), alpha);
        B.quad(d3(cx, cy, 0) - a * 0.5 - b * 0.5, a, b, (i % 2) ? M_WHITE : M_GOLD, 1.5);
    }
    // --- equipment boxes scattered over the truss, randomly oriented
    for (int i = 0; i < 160; ++i) {
        double x = r.range(-880.0, 880.0);
        double y = r.range(-13.0, 13.0), z = r.range(-12.0, 12.0);
        double sx = r.range(1.5, 6.0), sy = r.range(1.5, 6.0), sz = r.range(1.5, 5.0);
        double ax = r.range(0, 360), ay = r.range(0, 360), az = r.range(0, 360);
        auto R = [&](D3 v) { return rot_z(rot_y(rot_x(v, ax), ay), az); };
        static const int mats[8] = {M_GOLD, M_RED, M_GREEN, M_DARK, M_FOIL, M_GLASS, M_WHITE, M_BRUSHED};
        B.box(d3(x, y, z), R(d3(sx, 0, 0)), R(d3(0, sy, 0)), R(d3(0, 0, sz)), mats[i % 8], 3.0);
    }
    // --- dishes and docking rings
    for (int i = 0; i < 12; ++i) {
        D3 c = d3(r.range(-700.0, 700.0), r.range(-40.0, 40.0), r.range(8.0, 16.0) * ((i & 1) ? 1 : -1));
        D3 n = normd(d3(r.range(-0.5, 0.5), r.range(-0.5, 0.5), 1.0));
        Patch q{DISC, c, n, d3(r.range(4.0, 8.0), 0.08, 0), 0, B.log_density() * 2.0, (i % 2) ? M_STEEL : M_WHITE, true};
        q.area = M_PI * q.b.x * q.b.x;
        B.patches.push_back(q);
    }
}

struct Tri { D3 v[3]; int mat; };

// Tessellate every patch with a density multiplier `k`; returns total triangle count; when
// `out` is non-null the triangles are emitted.
size_t tessellate(const std::vector<Patch>& patches, double k, uint32_t seed, std::vector<Tri>* out) {
    size_t total = 0;
    for (size_t pi = 0; pi < patches.size(); ++pi) {
        const Patch& p = patches[pi];
        double lu, lv; patch_extent(p, lu, lv);
        double cells = std::max(1.0, 0.5 * k * p.density * p.area);   // 2 triangles per cell
        double aspect = (lv > 0 && lu > 0) ? lu / lv : 1.0;
        int nu = std::max(p.wrap_u ? 3 : 1, (int)std::floor(std::sqrt(cells * aspect) + 0.5));
        int nv = std::max(1, (int)std::floor(cells / nu + 0.5));
        total += (size_t)nu * nv * 2;
        if (!out) continue;
        const double cell = std::min(lu / nu, lv / nv);
        const double jit = 0.02 * cell;          // vertex displacement, well below a cell
        auto vert = [&](int i, int j) {
            int iw = p.wrap_u ? (i % nu) : i;
            D3 q = eval(p, (double)iw / nu, (double)j / nv);
            q.x += jit * (hash01(seed, pi, (uint64_t)iw * 131071u + j, 1) - 0.5);
            q.y += jit * (hash01(seed, pi, (uint64_t)iw * 131071u + j, 2) - 0.5);
            q.z += jit * (hash01(seed, pi, (uint64_t)iw * 131071u + j, 3) - 0.5);
            return q;
        };
        for (int j = 0; j < nv; ++j)
            for (int i = 0; i < nu; ++i) {
                D3 a = vert(i, j), b = vert(i + 1, j), c = vert(i + 1, j + 1), d = vert(i, j + 1);
                if ((i + j) & 1) {
                    out->push_back(Tri{{a, b, c}, p.mat});
                    out->push_back(Tri{{a, c, d}, p.mat});
                } else {
                    out->push_back(Tri{{a, b, d}, p.mat});
                    out->push_back(Tri{{b, c, d}, p.mat});
                }
            }
    }
    return total;
}

inline DsrtVec3 tof(D3 a) { DsrtVec3 r; r.x = (float)a.x; r.y = (float)a.y; r.z = (float)a.z; return r; }

// Flat normal with the reference's float arithmetic: unit_vector(cross(v1-v0, v2-v0)).
inline DsrtVec3 face_normal(DsrtVec3 v0, DsrtVec3 v1, DsrtVec3 v2) {
    float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z;
    float bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
    float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    float len = sqrtf(cx * cx + cy * cy + cz * cz);
    float inv = 1.0f / len;
    DsrtVec3 n; n.x = cx * inv; n.y = cy * inv; n.z = cz * inv;
    return n;
}

}  // namespace

extern "C" int dsrt_synth_mesh(int ntris, uint32_t seed, int diffuse_only,
                               DsrtTriangle* tris, DsrtMaterial* mats, int max_mats) {
    if (ntris < 0 || (ntris > 0 && !tris)) return DSRT_EINVAL;
    if (mats) {
        if (max_mats < M_COUNT) return DSRT_EINVAL;
        fill_materials(mats, diffuse_only);
    }
    if (ntris == 0) return M_COUNT;

    Builder B(seed);
    build_station(B);

    // Largest density multiplier whose tessellation does not exceed ntris.
    double lo = 0.0, hi = 1.0;
    while (tessellate(B.patches, hi, seed, nullptr) <= (size_t)ntris && hi < 1e12) hi *= 2.0;
    for (int it = 0; it < 60; ++it) {
        double mid = 0.5 * (lo + hi);
        if (tessellate(B.patches, mid, seed, nullptr) <= (size_t)ntris) lo = mid; else hi = mid;
    }
    std::vector<Tri> T;
    T.reserve((size_t)ntris + 16);
    size_t base = tessellate(B.patches, lo, seed, &T);
    if (base > (size_t)ntris) {           // tiny requests: fewer triangles than the coarsest tessellation
        // keep an evenly spaced subset
        std::vector<Tri> S;
        S.reserve(ntris);
        for (int i = 0; i < ntris; ++i) S.push_back(T[(size_t)((double)i * base / ntris)]);
        T.swap(S);
        base = T.size();
    }
    // Make up the deficit by bisecting evenly spaced triangles across edge v1-v2.
    size_t deficit = (size_t)ntris - base;
    while (deficit > 0) {
        size_t n0 = T.size();
        size_t take = std::min(deficit, n0);
        double step = (double)n0 / (double)take;
        for (size_t k = 0; k < take; ++k) {
            size_t i = (size_t)(k * step);
            Tri t = T[i];
            D3 m = (t.v[1] + t.v[2]) * 0.5;
            T[i] = Tri{{t.v[0], t.v[1], m}, t.mat};
            T.push_back(Tri{{t.v[0], m, t.v[2]}, t.mat});
        }
        deficit -= take;
    }

    // Sub-degree global rotation (keeps z within the +-25 envelope) and conversion to float.
    for (int i = 0; i < ntris; ++i) {
        DsrtTriangle o;
        std::memset(&o, 0, sizeof(o));
        DsrtVec3 v[3];
        for (int k = 0; k < 3; ++k) {
            D3 p = rot_z(rot_y(rot_x(T[i].v[k], 0.21), 0.13), 0.41);
            v[k] = tof(p);
        }
        o.v0 = v[0]; o.v1 = v[1]; o.v2 = v[2];
        DsrtVec3 n = face_normal(v[0], v[1], v[2]);
        o.n0 = n; o.n1 = n; o.n2 = n;
        o.material_id = T[i].mat;
        o.albedo_tex = -1;
        tris[i] = o;
    }
    return M_COUNT;
}
