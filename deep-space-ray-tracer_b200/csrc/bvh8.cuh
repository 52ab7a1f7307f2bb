// bvh8.cuh -- the compressed 8-wide BVH: memory layout and per-thread traversal.
//
// Replaces the reference's 40-byte binary GPUBVHNode (inc/gpu_scene.h:77-85) and its
// traversal bvh_hit_closest / bbox_hit / hit_triangle_index
// (/root/reference/src/gpu_render.cu:387-473, 285-315, 322-380).
//
// Node (96 bytes = six 16-byte words, fetched with six 128-bit loads):
//   w0: float p.x, p.y, p.z            quantisation origin (node box minimum, padded)
//       u32   ex | ey<<8 | ez<<16 | imask<<24     biased exponents of the grid step per axis,
//                                                  imask bit s = child slot s is an inner node
//   w1: u32 child_base   index of the first inner child (inner children are contiguous,
//                        ordered by slot)
//       u32 tri_base     index of the first triangle record of this node's leaf children
//       u8  meta[8]      slot s: 0 = empty; inner: 0x20 | (24+s); leaf: unary(count)<<5 | offset
//                        (the builder's description of the slot; the traversal reads w5 instead)
//   w2: u8 qlo_x[8], qlo_y[8]
//   w3: u8 qlo_z[8], qhi_x[8]
//   w4: u8 qhi_y[8], qhi_z[8]
//   w5: bf16 tri_weight[8]   slot s: the leaf child's triangle bits unary(count) << offset as a
//                            number (1, 3 or 7 times 2^offset -- two mantissa bits, so exact in
//                            bfloat16); 0 for inner and empty slots
// Child box = p + q * 2^(e-127) per axis, conservative by construction (bvh_build.cu).
// Children sit in slots chosen from their octant relative to the node centre, so visiting the
// hit inner children in decreasing `slot ^ octant-inverse` approximates front-to-back order
// without sorting (Ylitie, Karras, Laine: "Efficient incoherent ray traversal on GPUs
// through compressed wide BVHs", HPG 2017 -- algorithm restated, code written here).
//
// Hit masks are summed, not shifted together: the ALU pipe (PRMT, FMNMX, LOP3, SHF, SEL) is what
// bounds this traversal on B200 (profiles/r01_n: ALU 70 % busy, FMA 24 %, top stall
// math_pipe_throttle) and assembling `bits << index` per child cost ~8 ALU instructions for each
// of the eight slots.  Instead every hit child adds two floats on the FMA pipe under its hit
// predicate: 2^slot into an accumulator that starts at 2^23 (its low mantissa byte is then the
// 8-bit slot mask) and its tri_weight into a second one (sums of disjoint bits below 2^24 are
// exact).  The slot mask, restricted to inner children, goes through a 2 KB shared-memory table
// that applies the `slot ^ octant-inverse` permutation.
//
// Triangle record (48 bytes, three 128-bit loads), stored in leaf order:
//   a: v0.xyz, bits(index in the uploaded triangle array)
//   b: e1.xyz = v1 - v0 (one fp32 subtraction, as the reference computes it per test), 0
//   c: e2.xyz = v2 - v0, 0
// The accept test below is arithmetically identical to hit_triangle_index
// (gpu_render.cu:336-353) as compiled for sm_100a; see exact_math.cuh.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "exact_math.cuh"

namespace dsrt {

constexpr int kNodeWords = 6;          // uint4 words per node
constexpr int kTriWords = 3;           // float4 words per triangle record
constexpr int kMaxLeafTris = 3;
#ifndef DSRT_SMEM_STACK
#define DSRT_SMEM_STACK 8
#endif
constexpr int kSmemStack = DSRT_SMEM_STACK;          // stack entries per thread kept in shared memory
constexpr int kLocalStack = 40;        // overflow entries per thread in local memory
#ifndef DSRT_TRI_PER_STEP
#define DSRT_TRI_PER_STEP 2            // 0: a step drains the triangles its node visit exposed; n > 0: at most n per step (step_bounded).
                                       // Measured on B200 (profiles/r02_ab_variants.txt), C2 traversal ms: drain 8.71, n = 1: 8.84,
                                       // n = 2: 8.12, n = 3: 8.58; C5: 46.2 / 45.5 / 43.1 / 45.5
#endif
#ifndef DSRT_TRI_ROUND_MIN
#define DSRT_TRI_ROUND_MIN 0           // k > 0 (with DSRT_TRI_PER_STEP = n): the warp runs a first round of one test per lane
#endif                                 // and further rounds (up to n) only while at least k lanes still have a triangle pending
#ifndef DSRT_PREFETCH_NEXT
#define DSRT_PREFETCH_NEXT 0           // 1: visit() prefetches the node of the next visit into L1 (see there)
#endif

struct TravRay {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
};

struct TravHit {
    float t, u, v;
    int tri;       // position in the leaf-ordered record array, -1 = miss
};

__device__ __forceinline__ float safe_rcp_dir(float d) {
    // Box test only: keep the slab arithmetic finite for axis-parallel rays.  The triangle
    // test always sees the true direction.  One MUFU.RCP (1 ulp) instead of the IEEE reciprocal:
    // a relative error of 2^-23 in all plane distances of an axis is far inside what the box padding
    // (2^-19 of the scene size) and the widened far planes (2^-20) absorb.
    const float tiny = 1e-18f;
    const float a = fabsf(d) < tiny ? copysignf(tiny, d) : d;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

// What init() derives from a direction: {1/dx, 1/dy, 1/dz (safe_rcp_dir), bits(7 - octant)}.
__device__ __forceinline__ float4 prepare_dir(float dx, float dy, float dz) {
    const float ix = safe_rcp_dir(dx), iy = safe_rcp_dir(dy), iz = safe_rcp_dir(dz);
    const uint32_t oct = (ix < 0.0f ? 1u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 4u : 0u);
    return make_float4(ix, iy, iz, __uint_as_float(7u - oct));
}

// Per-thread traversal stack: the first kSmemStack entries live in shared memory laid out
// [entry][thread] (conflict-free 64-bit accesses across a warp), the rest in local memory.
struct TravStack {
    uint2* smem;        // &block_stack[threadIdx.x]; stride blockDim.x between entries
    int stride;
    uint2 local[kLocalStack];
    int sp;
    __device__ __forceinline__ void push(uint2 v) {
        if (sp < kSmemStack) smem[sp * stride] = v;
        else if (sp < kSmemStack + kLocalStack) local[sp - kSmemStack] = v;
        ++sp;   // deeper than smem+local cannot happen for trees of depth <= 48 (checked at build)
    }
    __device__ __forceinline__ uint2 pop() {
        --sp;
        return sp < kSmemStack ? smem[sp * stride] : local[sp - kSmemStack];
    }
};

// Moeller-Trumbore with the reference's exact operation order.  Returns true and updates
// (t,u,v) when the hit is accepted in the closed interval [tmin,tmax].
__device__ __forceinline__ bool tri_test(const float4 ra, const float4 rb, const float4 rc,
                                         V3 orig, V3 dir, float tmin, float tmax,
                                         float& t_out, float& u_out, float& v_out) {
    const V3 v0 = v3(ra.x, ra.y, ra.z);
    const V3 e1 = v3(rb.x, rb.y, rb.z);
    const V3 e2 = v3(rc.x, rc.y, rc.z);
    const V3 pvec = vcross(dir, e2);
    const float det = vdot(e1, pvec);
    // The |det| and u rejections share one branch: both need nothing but arithmetic, and a
    // separate early-out on det lets the compiler sink the v0 load below it, which serialises two
    // memory latencies per test (profiles/r01_c).  Same accept set as the reference's two tests.
    const float inv = __frcp_rn(det);
    const V3 tvec = vsub(orig, v0);
    const float u = __fmul_rn(vdot(tvec, pvec), inv);
    if (fabsf(det) < 1e-8f || u < 0.0f || u > 1.0f) return false;
    const V3 qvec = vcross(tvec, e1);
    const float v = __fmul_rn(vdot(dir, qvec), inv);
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    const float t = __fmul_rn(vdot(e2, qvec), inv);
    if (t < tmin || t > tmax) return false;
    t_out = t; u_out = u; v_out = v;
    return true;
}

// The three 128-bit loads of a triangle record, issued together.  Written as volatile asm
// because the compiler otherwise sinks the v0 load below the |det| early-out of tri_test, which
// puts a second full memory latency into every triangle test (profiles/r01_c: the triangle loop
// is 21 % of the traversal instructions but 41 % of its stall samples).
__device__ __forceinline__ void load_tri_record(const float4* tp, float4& ra, float4& rb, float4& rc) {
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(ra.x), "=f"(ra.y), "=f"(ra.z), "=f"(ra.w) : "l"(tp));
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4+16];"
                 : "=f"(rb.x), "=f"(rb.y), "=f"(rb.z), "=f"(rb.w) : "l"(tp));
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4+32];"
                 : "=f"(rc.x), "=f"(rc.y), "=f"(rc.z), "=f"(rc.w) : "l"(tp));
}

// Byte j of a packed plane word as the float 1 + q * 2^-15: one PRMT drops the byte into
// mantissa bits 8..15 of 1.0f (0x3F800000).  `one` must live in a register so that the PRMT
// selector can be the instruction's immediate; the caller gets it from memory the compiler
// cannot see through.  (profiles/r01_c: I2F.U8 is a quarter-rate XU instruction, and the earlier
// PRMT + FADD(-2^23) form cost two instructions per plane plus a register move for every selector.)
template <int J>
__device__ __forceinline__ float plane_byte(uint32_t w, uint32_t one) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(one), "n"(0x7604 | (J << 4)));
    return __uint_as_float(r);
}

// Slab constants of one axis for the folded form  t = fma(1 + q*2^-15, A, B):
//   A  = 2^15 * step * (1/d)               (exact: a power-of-two scaling of 1/d)
//   B0 = b - A * (1 + s*2^-22)             near planes, s = sign(A)
//   B1 = b - A * (1 - s*2^-22)             far planes,  b = (p - o) * (1/d)
// Computed as (b - A) -+ |A| 2^-22 -- the absolute value is an operand modifier, so no select on
// the ray's sign is needed (selects and compares run on the ALU pipe that bounds this loop).
// B carries a rounding error of at most 2^-23 (|b| + |A|): the |A| part is 2^-8 of one
// quantisation step in t, which the 2^-22 |A| bias (2^-7 of a step, towards "box larger")
// over-covers; the |b| part (2^-23 of the distance to the node, against a box padding of 2^-19 of
// the scene size) is of the size the builder's box padding absorbs.
struct SlabAxis { float A, B0, A1, B1; };
__device__ __forceinline__ SlabAxis slab_axis(uint32_t ebyte, float p, float o, float id) {
    constexpr float kWiden = 1.0f + 9.5367431640625e-07f;      // far planes: 2^-20 relative
    constexpr float kBias = 2.384185791015625e-07f;            // 2^-22
    SlabAxis s;
    s.A = __uint_as_float((ebyte << 23) + 0x07800000u) * id;
    const float bA = (p - o) * id - s.A;
    s.B0 = fmaf(fabsf(s.A), -kBias, bA);
    s.B1 = fmaf(fabsf(s.A), kBias, bA) * kWiden;
    s.A1 = s.A * kWiden;
    return s;
}

// Table of the octant permutation: entry [o][m] = the 8-bit mask m with bit s moved to bit s ^ o.
// 2 KB per block, filled by the whole block before its first traversal step.
constexpr int kPermTableBytes = 8 * 256;
__device__ __forceinline__ void fill_perm_table(uint8_t* table, int nthreads) {
    for (int i = threadIdx.x; i < kPermTableBytes; i += nthreads) {
        const uint32_t o = (uint32_t)i >> 8, m = (uint32_t)i & 0xFFu;
        uint32_t r = 0u;
#pragma unroll
        for (uint32_t sl = 0; sl < 8u; ++sl)
            if ((m >> sl) & 1u) r |= 1u << (sl ^ o);
        table[i] = (uint8_t)r;
    }
    __syncthreads();
}

// Intersects the eight child boxes of one node.  Returns the hit inner children as an 8-bit mask in
// slot order (still to be permuted into traversal order) and sets `tri_hits` to the 24-bit mask of
// the triangles of the hit leaf children.
__device__ __forceinline__ uint32_t node_test(const uint4* __restrict__ nodes, uint32_t node_index,
                                              float ox, float oy, float oz,
                                              float idx, float idy, float idz,
                                              float tmin, float tmax, uint32_t one,
                                              uint32_t& child_base, uint32_t& tri_base, uint32_t& imask,
                                              uint32_t& tri_hits) {
    const uint4* np = nodes + (size_t)node_index * kNodeWords;
    const uint4 w0 = __ldg(np + 0);
    const uint4 w1 = __ldg(np + 1);
    const uint4 w2 = __ldg(np + 2);
    const uint4 w3 = __ldg(np + 3);
    const uint4 w4 = __ldg(np + 4);
    const uint4 w5 = __ldg(np + 5);

    const uint32_t ew = w0.w;
    imask = ew >> 24;
    child_base = w1.x;
    tri_base = w1.y;

    const SlabAxis X = slab_axis(ew & 0xFFu, __uint_as_float(w0.x), ox, idx);
    const SlabAxis Y = slab_axis((ew >> 8) & 0xFFu, __uint_as_float(w0.y), oy, idy);
    const SlabAxis Z = slab_axis((ew >> 16) & 0xFFu, __uint_as_float(w0.z), oz, idz);

    float slot_acc = 8388608.0f;       // 2^23: the slots hit end up in the low mantissa byte
    float tri_acc = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t qlox = h ? w2.y : w2.x, qloy = h ? w2.w : w2.z;
        const uint32_t qloz = h ? w3.y : w3.x, qhix = h ? w3.w : w3.z;
        const uint32_t qhiy = h ? w4.y : w4.x, qhiz = h ? w4.w : w4.z;
        const uint32_t nx = idx < 0.0f ? qhix : qlox, fx = idx < 0.0f ? qlox : qhix;
        const uint32_t ny = idy < 0.0f ? qhiy : qloy, fy = idy < 0.0f ? qloy : qhiy;
        const uint32_t nz = idz < 0.0f ? qhiz : qloz, fz = idz < 0.0f ? qloz : qhiz;
        const uint32_t wa = h ? w5.z : w5.x, wb = h ? w5.w : w5.y;       // weights of slots 4h..4h+1, 4h+2..4h+3
#define DSRT_CHILD(J, WEIGHT)                                                              \
        {                                                                                  \
            const float t0x = fmaf(plane_byte<J>(nx, one), X.A, X.B0);                     \
            const float t0y = fmaf(plane_byte<J>(ny, one), Y.A, Y.B0);                     \
            const float t0z = fmaf(plane_byte<J>(nz, one), Z.A, Z.B0);                     \
            const float t1x = fmaf(plane_byte<J>(fx, one), X.A1, X.B1);                    \
            const float t1y = fmaf(plane_byte<J>(fy, one), Y.A1, Y.B1);                    \
            const float t1z = fmaf(plane_byte<J>(fz, one), Z.A1, Z.B1);                    \
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));                   \
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax));                   \
            if (cmin <= cmax) {                                                            \
                slot_acc += (float)(1 << (4 * h + J));                                     \
                tri_acc += __uint_as_float(WEIGHT);                                        \
            }                                                                              \
        }
        DSRT_CHILD(0, wa << 16) DSRT_CHILD(1, wa & 0xFFFF0000u) DSRT_CHILD(2, wb << 16) DSRT_CHILD(3, wb & 0xFFFF0000u)
#undef DSRT_CHILD
    }
    tri_hits = __float2uint_rz(tri_acc);
    return __float_as_uint(slot_acc) & imask;
}

// Traversal state of one ray, kept in registers by a persistent thread.  `step` advances the
// CWBVH loop by one node visit plus the triangles that visit exposed and returns true when
// the ray is finished (stack empty, or first accepted triangle for an any-hit ray).
// The triangle loop runs at ~4 of 32 lanes (profiles/r01_n: a node visit exposes 0.3 triangles on
// average) and is 20 % of the issued instructions; capping it at one or two tests per step, so that
// the other lanes go on visiting nodes, was measured 15-40 % SLOWER on B200 (C2 traversal 14.5 ->
// 16.8 ms at two per step, 20.2 ms at one): draining keeps t_max tight for the next node visit.
// The wait for a triangle record is the hottest single stall of the traversal (13 % of k_trace's
// samples); prefetching the later records of a group into L1 and fetching two records per trip
// were both measured slower as well (+3 %: the loop runs at the 64-register cap).
// TRACK = true (parity hook only) also records the node whose leaf child held the accepted triangle.
template <bool TRACK>
struct LaneTravT {
    float ox, oy, oz, dx, dy, dz;
    float idx, idy, idz;
    float tmin, tmax;             // tmax shrinks to the closest accepted hit
    float hu, hv;
    int htri;                     // position in the leaf-ordered record array, -1 = none, -2 - i = sphere i
    uint32_t hnode;               // TRACK only: the node visited when htri was accepted
    uint32_t tnode;               // TRACK only: the node whose triangles are pending in tpend
    uint2 tpend;                  // DSRT_TRI_PER_STEP > 0: triangles of the last node visit that still wait for their test
    uint32_t octinv;
    uint2 ngroup;
    uint32_t one;                 // 0x3F800000 held in a register (see plane_byte)
    const uint8_t* perm;          // the block's octant-permutation table (fill_perm_table), shared memory
    bool any_hit;

    __device__ __forceinline__ void init(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_,
                                         float tmin_, float tmax_, bool any_hit_, TravStack& stack) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_;
        idx = safe_rcp_dir(dx_); idy = safe_rcp_dir(dy_); idz = safe_rcp_dir(dz_);
        tmin = tmin_; tmax = tmax_;
        hu = 0.f; hv = 0.f; htri = -1;
        if (TRACK) hnode = 0xFFFFFFFFu;
        const uint32_t oct = (idx < 0.0f ? 1u : 0u) | (idy < 0.0f ? 2u : 0u) | (idz < 0.0f ? 4u : 0u);
        octinv = 7u - oct;
        ngroup = make_uint2(0u, 0x80000000u);      // the root as a one-bit node group
        tpend = make_uint2(0u, 0u);
        any_hit = any_hit_;
        stack.sp = 0;
    }

    // init() for a ray whose inverse direction and octant were prepared beforehand (prepare_dir) -- by a whole warp
    // for the rays it has staged, instead of by the few lanes that happen to take a new ray in a given iteration.
    __device__ __forceinline__ void init_prepared(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_, float4 inv,
                                                  float tmin_, float tmax_, bool any_hit_, TravStack& stack) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_;
        idx = inv.x; idy = inv.y; idz = inv.z;
        tmin = tmin_; tmax = tmax_;
        hu = 0.f; hv = 0.f; htri = -1;
        if (TRACK) hnode = 0xFFFFFFFFu;
        octinv = __float_as_uint(inv.w);
        ngroup = make_uint2(0u, 0x80000000u);
        tpend = make_uint2(0u, 0u);
        any_hit = any_hit_;
        stack.sp = 0;
    }

    __device__ __forceinline__ void set_hnode(uint32_t node) { if (TRACK) hnode = node; }

    __device__ __forceinline__ bool step(const uint4* __restrict__ nodes, const float4* __restrict__ tris,
                                         TravStack& stack) {
        uint2 tgroup;
        uint32_t cur_node = 0u;
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rel = __popc(ngroup.y & 0xFFu & ((1u << slot) - 1u));
            const uint32_t node_index = ngroup.x + rel;
            cur_node = node_index;
            if (ngroup.y > 0x00FFFFFFu) stack.push(ngroup);
            uint32_t child_base, tri_base, imask, tri_hits;
            const uint32_t slots = node_test(nodes, node_index, ox, oy, oz, idx, idy, idz, tmin, tmax,
                                             one, child_base, tri_base, imask, tri_hits);
            ngroup = make_uint2(child_base, ((uint32_t)perm[(octinv << 8) + slots] << 24) | imask);
            tgroup = make_uint2(tri_base, tri_hits);
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }
        const V3 orig = v3(ox, oy, oz), dir = v3(dx, dy, dz);
        while (tgroup.y != 0u) {
            const uint32_t bit = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << bit);
            const uint32_t ti = tgroup.x + bit;
            const float4* tp = tris + (size_t)ti * kTriWords;
            float4 ra, rb, rc;
            load_tri_record(tp, ra, rb, rc);
            float t, u, v;
            if (tri_test(ra, rb, rc, orig, dir, tmin, tmax, t, u, v)) {
                tmax = t; hu = u; hv = v; htri = (int)ti;
                if (TRACK) hnode = cur_node;
                if (any_hit) return true;
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (stack.sp == 0) return true;
            ngroup = stack.pop();
        }
        return false;
    }

    // step() with at most DSRT_TRI_PER_STEP triangle tests per call.  A node visit exposes 0.3 triangles per ray on
    // average but now and then several at once, and in step() the whole warp waits while one lane drains its three:
    // the triangle loop ran ~2.5 rounds per step at 3.6 of 32 lanes (profiles/r02_b_source_lines.txt).  Here a lane
    // tests a bounded number per call and KEEPS the rest (tpend); while it has triangles pending it does not visit
    // another node -- its t_max stays as tight as in step(), it just sits out the node test of the next call -- so
    // every call costs one node visit and one short round of triangle tests at three times the width.  Hits are
    // those of step() (the same triangles are tested against the same shrinking interval, in the same order).
    __device__ __forceinline__ bool step_bounded(const uint4* __restrict__ nodes, const float4* __restrict__ tris,
                                                 TravStack& stack, int max_tests) {
        if (tpend.y == 0u) {
            uint32_t cur_node;
            tpend = visit(nodes, stack, cur_node);
            if (TRACK) tnode = cur_node;
        }
        const V3 orig = v3(ox, oy, oz), dir = v3(dx, dy, dz);
        for (int k = 0; k < max_tests && tpend.y != 0u; ++k) {
            const uint32_t bit = 31u - __clz(tpend.y);
            tpend.y &= ~(1u << bit);
            const uint32_t ti = tpend.x + bit;
            float4 ra, rb, rc;
            load_tri_record(tris + (size_t)ti * kTriWords, ra, rb, rc);
            float t, u, v;
            if (tri_test(ra, rb, rc, orig, dir, tmin, tmax, t, u, v)) {
                tmax = t; hu = u; hv = v; htri = (int)ti;
                if (TRACK) hnode = tnode;
                if (any_hit) return true;
            }
        }
        if (tpend.y != 0u) return false;          // the rest of this node's triangles: next call
        return settle(stack, false);
    }

    // step_bounded() in pieces, for trace loops that decide per WARP how many rounds of triangle tests a step gets
    // (DSRT_TRI_ROUND_MIN): the node visit of a lane with nothing pending, and one triangle test of a lane with something.
    __device__ __forceinline__ void visit_pending(const uint4* __restrict__ nodes, TravStack& stack) {
        uint32_t cur_node;
        tpend = visit(nodes, stack, cur_node);
        if (TRACK) tnode = cur_node;
    }
    // Tests the first pending triangle; true when an any-hit ray found its occluder.
    __device__ __forceinline__ bool test_one(const float4* __restrict__ tris) {
        const uint32_t bit = 31u - __clz(tpend.y);
        tpend.y &= ~(1u << bit);
        const uint32_t ti = tpend.x + bit;
        float4 ra, rb, rc;
        load_tri_record(tris + (size_t)ti * kTriWords, ra, rb, rc);
        float t, u, v;
        if (tri_test(ra, rb, rc, v3(ox, oy, oz), v3(dx, dy, dz), tmin, tmax, t, u, v)) {
            tmax = t; hu = u; hv = v; htri = (int)ti;
            if (TRACK) hnode = tnode;
            if (any_hit) return true;
        }
        return false;
    }

    // ---- step() in three parts, for the trace loops that test triangles warp-cooperatively (pooled_tri_test) ----
    // visit(): the node half of step().  Returns the triangles this visit exposed {first record, 24-bit mask}.
    __device__ __forceinline__ uint2 visit(const uint4* __restrict__ nodes, TravStack& stack, uint32_t& cur_node) {
        uint2 tgroup;
        cur_node = 0u;
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rel = __popc(ngroup.y & 0xFFu & ((1u << slot) - 1u));
            const uint32_t node_index = ngroup.x + rel;
            cur_node = node_index;
            if (ngroup.y > 0x00FFFFFFu) stack.push(ngroup);
            uint32_t child_base, tri_base, imask, tri_hits;
            const uint32_t slots = node_test(nodes, node_index, ox, oy, oz, idx, idy, idz, tmin, tmax,
                                             one, child_base, tri_base, imask, tri_hits);
            ngroup = make_uint2(child_base, ((uint32_t)perm[(octinv << 8) + slots] << 24) | imask);
            tgroup = make_uint2(tri_base, tri_hits);
#if DSRT_PREFETCH_NEXT
            // the node the next visit will fetch is known now: start its two cache lines on their way while the
            // triangles of this visit are tested
            if (ngroup.y > 0x00FFFFFFu) {
                const uint32_t nbit = 31u - __clz(ngroup.y);
                const uint32_t nslot = (nbit - 24u) ^ octinv;
                const uint32_t nrel = __popc(imask & ((1u << nslot) - 1u));
                const char* np = reinterpret_cast<const char*>(nodes + (size_t)(child_base + nrel) * kNodeWords);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 80));
            }
#endif
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }
        return tgroup;
    }
    // The sequential triangle loop of step() on its own; true when an any-hit ray found its occluder.
    __device__ __forceinline__ bool test_own(const float4* __restrict__ tris, uint2 tgroup, uint32_t cur_node) {
        const V3 orig = v3(ox, oy, oz), dir = v3(dx, dy, dz);
        while (tgroup.y != 0u) {
            const uint32_t bit = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << bit);
            const uint32_t ti = tgroup.x + bit;
            float4 ra, rb, rc;
            load_tri_record(tris + (size_t)ti * kTriWords, ra, rb, rc);
            float t, u, v;
            if (tri_test(ra, rb, rc, orig, dir, tmin, tmax, t, u, v)) {
                tmax = t; hu = u; hv = v; htri = (int)ti;
                if (TRACK) hnode = cur_node;
                if (any_hit) return true;
            }
        }
        return false;
    }
    // settle(): the end of step(): true when the ray is finished.
    __device__ __forceinline__ bool settle(TravStack& stack, bool occluded) {
        if (occluded) return true;
        if (ngroup.y <= 0x00FFFFFFu) {
            if (stack.sp == 0) return true;
            ngroup = stack.pop();
        }
        return false;
    }

    // The brute-force sphere loop of the reference's scene_hit (gpu_render.cu:530-546), run once
    // the triangles are done: the interval's upper end is the closest triangle hit and it is
    // closed on both ends, so a sphere at exactly that distance replaces the triangle.
    // hit_sphere (gpu_render.cu:478-504) as compiled: c = fma(-r, r, dot(oc, oc)),
    // discriminant = fma(hb, hb, -rn(a * c)), IEEE sqrt and divisions.
    // Records: 2 x float4 per sphere {centre.xyz, radius | 1/radius, bits(material), 0, 0}.
    // A sphere hit is reported as htri = -2 - index.
    // `a` = dot(dir, dir) is loop-invariant and the reference's compiler hoists and contracts it
    // differently in its inlined copies: in render_kernel's closest-hit loop it is
    // fma(z,z, fma(y,y, rn(x*x))) (cuobjdump: FMUL x,x; FFMA y,y; FFMA z,z), in the shadow-ray
    // loop and in the stand-alone scene_hit the usual fma(z,z, fma(x,x, rn(y*y))).  On a
    // sphere's silhouette and terminator that last bit decides hit or miss, so the render path
    // passes x_first = true for its closest-hit rays.
#ifndef DSRT_SPHERES_INLINE
#define DSRT_SPHERES_INLINE __forceinline__
#endif
    __device__ DSRT_SPHERES_INLINE void spheres_pass(const float4* __restrict__ sph, int n, bool x_first) {
        if (any_hit && htri != -1) return;
        const V3 orig = v3(ox, oy, oz), dir = v3(dx, dy, dz);
        const float a = (x_first && !any_hit) ? __fmaf_rn(dir.z, dir.z, __fmaf_rn(dir.y, dir.y, __fmul_rn(dir.x, dir.x)))
                                              : vdot(dir, dir);
        for (int i = 0; i < n; ++i) {
            const float4 s0 = __ldg(&sph[2 * i]);
            const V3 oc = vsub(orig, v3(s0.x, s0.y, s0.z));
            const float half_b = vdot(oc, dir);
            const float c = __fmaf_rn(-s0.w, s0.w, vdot(oc, oc));
            const float disc = __fmaf_rn(half_b, half_b, -__fmul_rn(a, c));
            if (disc < 0.0f) continue;
            const float sq = __fsqrt_rn(disc);
            float root = __fdiv_rn(__fsub_rn(-half_b, sq), a);
            if (root < tmin || root > tmax) {
                root = __fdiv_rn(__fadd_rn(-half_b, sq), a);
                if (root < tmin || root > tmax) continue;
            }
            tmax = root; hu = 0.0f; hv = 0.0f; htri = -2 - i;
            if (any_hit) return;
        }
    }
};
using LaneTrav = LaneTravT<false>;

// ---------------------------------------------------------------------------------------
// Warp-cooperative triangle test.
//
// A node visit exposes 0.3 triangles per ray on average, so the per-lane triangle loop of step() runs at 3-5
// of 32 lanes for as many rounds as the busiest lane has triangles (profiles/r02_b: 30 % of the traversal
// kernels' instructions at 3.6 threads per instruction in k_trace, 5 in k_primary).  Here the triangles that
// the lanes of a warp exposed in one step are pooled in shared memory (an exclusive scan of the per-lane
// counts gives every owner its place), each lane then tests ONE triangle of the pool -- fetching the owner's
// ray with shuffles -- and writes {t, u, v, accepted} back; finally every owner folds its own results in the
// order the sequential loop would have tested them, with the same acceptance rule (not t > t_max of the
// moment), so the outcome is the sequential loop's bit for bit, equal-t ties included.
// ---------------------------------------------------------------------------------------
constexpr int kPoolCap = 32;             // triangles tested per round: one per lane
constexpr uint32_t kPoolIndexBits = 27;  // record positions below 2^27 (checked at upload); the owner lane sits above
struct TriPool {                         // one per warp
    uint32_t entry[kPoolCap];            // owner lane << 27 | record position
    float4 result[kPoolCap];             // t, u, v, accepted (1.0f / 0.0f)
};
#ifndef DSRT_POOL_MIN_COUNT
#define DSRT_POOL_MIN_COUNT 99           // pool only when some lane holds at least this many triangles; above 24 = never.
#endif                                   // MEASURED ON B200 AND SWITCHED OFF (profiles/r02_ab_variants.txt): C2 traversal 9.17 ms
                                         // with the per-lane loop, 9.67 ms pooled (threshold 2), 9.52 (3), 9.68 (always), all at 7
                                         // blocks per SM (the pooled loop needs 72 registers; at 64 it spills and takes 10.6 ms).
                                         // The scan, seven shuffles, two shared-memory round trips and three warp barriers per step
                                         // cost more than the 2.5 sequential rounds at ~4 lanes that they replace.

// Called by all 32 lanes in converged code; `tg` is what visit() returned (ignored for lanes without a ray).
// Returns true for an any-hit ray that found its occluder.
template <class Trav>
__device__ __forceinline__ bool pooled_tri_test(Trav& L, bool active, uint2 tg, uint32_t cur_node,
                                                const float4* __restrict__ tris, TriPool* pool, uint32_t lane) {
    uint32_t bits = active ? tg.y : 0u;
#if DSRT_POOL_MIN_COUNT > 24
    if (bits != 0u) return L.test_own(tris, make_uint2(tg.x, bits), cur_node);      // never pool (A/B builds)
    return false;
#endif
    if (__ballot_sync(0xffffffffu, __popc(bits) >= DSRT_POOL_MIN_COUNT) == 0u) {
        if (bits != 0u) return L.test_own(tris, make_uint2(tg.x, bits), cur_node);
        return false;
    }
    bool occluded = false;
    const uint32_t index_mask = (1u << kPoolIndexBits) - 1u;
    while (true) {
        const uint32_t cnt = __popc(bits);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += n;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0u) break;
        const uint32_t excl = incl - cnt;
        // owners list their triangles in the order the sequential loop tests them (highest bit first)
        uint32_t my_n = 0u;
        while (bits != 0u && excl + my_n < (uint32_t)kPoolCap) {
            const uint32_t bit = 31u - __clz(bits);
            bits &= ~(1u << bit);
            pool->entry[excl + my_n] = (lane << kPoolIndexBits) | (tg.x + bit);
            ++my_n;
        }
        __syncwarp();
        const uint32_t n_pool = min(total, (uint32_t)kPoolCap);
        const bool has = lane < n_pool;
        const uint32_t e = has ? pool->entry[lane] : 0u;
        const uint32_t owner = e >> kPoolIndexBits;
        const V3 orig = v3(__shfl_sync(0xffffffffu, L.ox, owner), __shfl_sync(0xffffffffu, L.oy, owner), __shfl_sync(0xffffffffu, L.oz, owner));
        const V3 dir = v3(__shfl_sync(0xffffffffu, L.dx, owner), __shfl_sync(0xffffffffu, L.dy, owner), __shfl_sync(0xffffffffu, L.dz, owner));
        const float tmin = __shfl_sync(0xffffffffu, L.tmin, owner);
        if (has) {
            float4 ra, rb, rc;
            load_tri_record(tris + (size_t)(e & index_mask) * kTriWords, ra, rb, rc);
            float t = 0.f, u = 0.f, v = 0.f;
            // the upper end of the interval is the owner's business (its t_max shrinks as it folds the results)
            const bool ok = tri_test(ra, rb, rc, orig, dir, tmin, __int_as_float(0x7f800000), t, u, v);
            pool->result[lane] = make_float4(t, u, v, ok ? 1.0f : 0.0f);
        }
        __syncwarp();
        for (uint32_t k = 0; k < my_n; ++k) {
            const float4 r = pool->result[excl + k];
            if (r.w != 0.0f && !(r.x > L.tmax)) {
                L.tmax = r.x; L.hu = r.y; L.hv = r.z;
                L.htri = (int)(pool->entry[excl + k] & index_mask);
                L.set_hnode(cur_node);
                if (L.any_hit) { occluded = true; bits = 0u; break; }
            }
        }
        __syncwarp();                     // the pool is rewritten by the next round
    }
    return occluded;
}

}  // namespace dsrt
