// bvh8.cuh -- the compressed 8-wide BVH: memory layout and per-thread traversal.
//
// Replaces the reference's 40-byte binary GPUBVHNode (inc/gpu_scene.h:77-85) and its
// traversal bvh_hit_closest / bbox_hit / hit_triangle_index
// (/root/reference/src/gpu_render.cu:387-473, 285-315, 322-380).
//
// Node (80 bytes = five 16-byte words, fetched with five 128-bit loads):
//   w0: float p.x, p.y, p.z            quantisation origin (node box minimum, padded)
//       u32   ex | ey<<8 | ez<<16 | imask<<24     biased exponents of the grid step per axis,
//                                                  imask bit s = child slot s is an inner node
//   w1: u32 child_base   index of the first inner child (inner children are contiguous,
//                        ordered by slot)
//       u32 tri_base     index of the first triangle record of this node's leaf children
//       u8  meta[8]      slot s: 0 = empty; inner: 0x20 | (24+s); leaf: unary(count)<<5 | offset
//   w2: u8 qlo_x[8], qlo_y[8]
//   w3: u8 qlo_z[8], qhi_x[8]
//   w4: u8 qhi_y[8], qhi_z[8]
// Child box = p + q * 2^(e-127) per axis, conservative by construction (bvh_build.cu).
// Children sit in slots chosen from their octant relative to the node centre, so visiting
// set bits of `(24+slot) ^ octant-inverse` from the top approximates front-to-back order
// without sorting (Ylitie, Karras, Laine: "Efficient incoherent ray traversal on GPUs
// through compressed wide BVHs", HPG 2017 -- algorithm restated, code written here).
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

constexpr int kNodeWords = 5;          // uint4 words per node
constexpr int kTriWords = 3;           // float4 words per triangle record
constexpr int kMaxLeafTris = 3;
constexpr int kSmemStack = 8;          // stack entries per thread kept in shared memory
constexpr int kLocalStack = 40;        // overflow entries per thread in local memory

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
    // test always sees the true direction.
    const float tiny = 1e-18f;
    float a = fabsf(d) < tiny ? copysignf(tiny, d) : d;
    return 1.0f / a;
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
    if (fabsf(det) < 1e-8f) return false;
    const float inv = __frcp_rn(det);
    const V3 tvec = vsub(orig, v0);
    const float u = __fmul_rn(vdot(tvec, pvec), inv);
    if (u < 0.0f || u > 1.0f) return false;
    const V3 qvec = vcross(tvec, e1);
    const float v = __fmul_rn(vdot(dir, qvec), inv);
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    const float t = __fmul_rn(vdot(e2, qvec), inv);
    if (t < tmin || t > tmax) return false;
    t_out = t; u_out = u; v_out = v;
    return true;
}

__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    // each byte: 0x80 set -> 0xFF, else 0x00
    return ((x >> 7) & 0x01010101u) * 0xFFu;
}

// Byte j of a packed word as a float.  I2F.U8 runs on the XU pipe (16 lanes/clk/SM) and 48 of
// them per node visit saturate it (profiles/r01_a: XU 81 % busy).  The alternative builds the
// float 2^23 + byte with one PRMT (ALU pipe) and subtracts 2^23 with one FADD (FMA pipe):
// exact, and spread over the two wide pipes.  kI2fMask picks, per plane set (near x,y,z, far
// x,y,z = bits 0..5), which conversion is used so the three pipes can be balanced.
template <bool USE_I2F>
__device__ __forceinline__ float byte_to_float(uint32_t w, int j) {
    if (USE_I2F) return (float)((w >> (8 * j)) & 0xFFu);
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | (uint32_t)j)) - 8388608.0f;
}

#ifndef DSRT_I2F_MASK
#define DSRT_I2F_MASK 0x09      // near-x and far-x through I2F, the other four through PRMT+FADD
#endif
constexpr int kI2fMask = DSRT_I2F_MASK;

// Intersects the eight child boxes of one node; returns the hit mask (bits 24..31 inner
// children in traversal-priority order, bits 0..23 triangles).
__device__ __forceinline__ uint32_t node_test(const uint4* __restrict__ nodes, uint32_t node_index,
                                              float ox, float oy, float oz,
                                              float idx, float idy, float idz,
                                              float tmin, float tmax, uint32_t octinv4,
                                              uint32_t& child_base, uint32_t& tri_base, uint32_t& imask) {
    const uint4* np = nodes + (size_t)node_index * kNodeWords;
    const uint4 w0 = __ldg(np + 0);
    const uint4 w1 = __ldg(np + 1);
    const uint4 w2 = __ldg(np + 2);
    const uint4 w3 = __ldg(np + 3);
    const uint4 w4 = __ldg(np + 4);

    const float px = __uint_as_float(w0.x), py = __uint_as_float(w0.y), pz = __uint_as_float(w0.z);
    const uint32_t ew = w0.w;
    imask = ew >> 24;
    child_base = w1.x;
    tri_base = w1.y;

    const float ax = __uint_as_float((ew & 0xFFu) << 23) * idx;
    const float ay = __uint_as_float(((ew >> 8) & 0xFFu) << 23) * idy;
    const float az = __uint_as_float(((ew >> 16) & 0xFFu) << 23) * idz;
    const float bx = (px - ox) * idx;
    const float by = (py - oy) * idy;
    const float bz = (pz - oz) * idz;
    // far planes widened by 2^-20 (relative): the slab test must never reject what the exact
    // triangle test accepts, whatever the distance between ray origin and node.
    constexpr float kWiden = 1.0f + 9.5367431640625e-07f;
    const float ax1 = ax * kWiden, ay1 = ay * kWiden, az1 = az * kWiden;
    const float bx1 = bx * kWiden, by1 = by * kWiden, bz1 = bz * kWiden;

    uint32_t hitmask = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t meta4 = h ? w1.w : w1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;

        const uint32_t qlox = h ? w2.y : w2.x, qloy = h ? w2.w : w2.z;
        const uint32_t qloz = h ? w3.y : w3.x, qhix = h ? w3.w : w3.z;
        const uint32_t qhiy = h ? w4.y : w4.x, qhiz = h ? w4.w : w4.z;
        const uint32_t nx = idx < 0.0f ? qhix : qlox, fx = idx < 0.0f ? qlox : qhix;
        const uint32_t ny = idy < 0.0f ? qhiy : qloy, fy = idy < 0.0f ? qloy : qhiy;
        const uint32_t nz = idz < 0.0f ? qhiz : qloz, fz = idz < 0.0f ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sh = 8 * j;
            const float t0x = fmaf(byte_to_float<(kI2fMask & 1) != 0>(nx, j), ax, bx);
            const float t0y = fmaf(byte_to_float<(kI2fMask & 2) != 0>(ny, j), ay, by);
            const float t0z = fmaf(byte_to_float<(kI2fMask & 4) != 0>(nz, j), az, bz);
            const float t1x = fmaf(byte_to_float<(kI2fMask & 8) != 0>(fx, j), ax1, bx1);
            const float t1y = fmaf(byte_to_float<(kI2fMask & 16) != 0>(fy, j), ay1, by1);
            const float t1z = fmaf(byte_to_float<(kI2fMask & 32) != 0>(fz, j), az1, bz1);
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            if (cmin <= cmax) {
                const uint32_t bits = (child_bits4 >> sh) & 0xFFu;
                const uint32_t idxb = (bit_index4 >> sh) & 0xFFu;
                hitmask |= bits << idxb;
            }
        }
    }
    return hitmask;
}

// Traversal state of one ray, kept in registers by a persistent thread.  `step` advances the
// CWBVH loop by one node visit plus the triangles that visit exposed and returns true when
// the ray is finished (stack empty, or first accepted triangle for an any-hit ray).
struct LaneTrav {
    float ox, oy, oz, dx, dy, dz;
    float idx, idy, idz;
    float tmin, tmax;             // tmax shrinks to the closest accepted hit
    float hu, hv;
    int htri;                     // position in the leaf-ordered record array, -1 = none
    uint32_t octinv;
    uint2 ngroup;
    bool any_hit;

    __device__ __forceinline__ void init(float ox_, float oy_, float oz_, float dx_, float dy_, float dz_,
                                         float tmin_, float tmax_, bool any_hit_, TravStack& stack) {
        ox = ox_; oy = oy_; oz = oz_; dx = dx_; dy = dy_; dz = dz_;
        idx = safe_rcp_dir(dx_); idy = safe_rcp_dir(dy_); idz = safe_rcp_dir(dz_);
        tmin = tmin_; tmax = tmax_;
        hu = 0.f; hv = 0.f; htri = -1;
        const uint32_t oct = (idx < 0.0f ? 1u : 0u) | (idy < 0.0f ? 2u : 0u) | (idz < 0.0f ? 4u : 0u);
        octinv = 7u - oct;
        ngroup = make_uint2(0u, 0x80000000u);      // the root as a one-bit node group
        any_hit = any_hit_;
        stack.sp = 0;
    }

    __device__ __forceinline__ bool step(const uint4* __restrict__ nodes, const float4* __restrict__ tris,
                                         TravStack& stack) {
        uint2 tgroup;
        if (ngroup.y > 0x00FFFFFFu) {
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rel = __popc(ngroup.y & 0xFFu & ((1u << slot) - 1u));
            const uint32_t node_index = ngroup.x + rel;
            if (ngroup.y > 0x00FFFFFFu) stack.push(ngroup);
            uint32_t child_base, tri_base, imask;
            const uint32_t hm = node_test(nodes, node_index, ox, oy, oz, idx, idy, idz, tmin, tmax,
                                          octinv * 0x01010101u, child_base, tri_base, imask);
            ngroup = make_uint2(child_base, (hm & 0xFF000000u) | imask);
            tgroup = make_uint2(tri_base, hm & 0x00FFFFFFu);
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
            const float4 ra = __ldg(tp + 0);
            const float4 rb = __ldg(tp + 1);
            const float4 rc = __ldg(tp + 2);
            float t, u, v;
            if (tri_test(ra, rb, rc, orig, dir, tmin, tmax, t, u, v)) {
                tmax = t; hu = u; hv = v; htri = (int)ti;
                if (any_hit) return true;
            }
        }
        if (ngroup.y <= 0x00FFFFFFu) {
            if (stack.sp == 0) return true;
            ngroup = stack.pop();
        }
        return false;
    }
};

}  // namespace dsrt
