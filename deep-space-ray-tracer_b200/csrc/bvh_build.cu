// bvh_build.cu -- GPU construction of the compressed 8-wide BVH.
//
// Replaces the reference's per-frame, single-threaded CPU builder
// (/root/reference/src/gpu_scene_builder.cpp:343-459: recursive median split, leaves <= 4)
// with a build that runs once per mesh on the device:
//
//   1. triangle bounds + scene bounds            (k_tri_bounds)
//   2. 63-bit Morton codes of box centres        (k_morton)
//   3. radix sort (cub::DeviceRadixSort)
//   4. binary tree over the sorted triangles, one of
//        a. PLOC -- parallel locally-ordered clustering (Meister & Bittner 2018, algorithm
//           restated): every cluster looks +-kPlocRadius positions along the Morton order for the
//           partner that minimises the merged surface area, mutual nearest neighbours merge,
//           the cluster list is compacted, repeat until one cluster is left
//           (k_ploc_nearest, k_ploc_mark, cub::DeviceScan, k_ploc_merge) -- the default;
//        b. the binary radix tree of Karras 2012 (k_hierarchy) -- DSRT_BVH_BUILDER=lbvh
//   5. bottom-up pass with atomic flags: boxes, triangle counts and the dynamic programme below (k_refit)
//   6. top-down collapse, one level per launch:  (k_collapse_level)
//        each binary node expands into <= 8 children along the cost-optimal cut found by a
//        bottom-up dynamic programme over the surface-area heuristic (in k_refit),
//        subtrees of <= 3 triangles may become leaf children,
//        children are placed in octant-matched slots,
//        child boxes are quantised to 8 bits against the node box (conservatively),
//        nodes are emitted breadth first, triangle records in leaf order.
//
// The traversal that consumes this layout is in bvh8.cuh.
#include "bvh_build.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "bvh8.cuh"

namespace dsrt {

namespace {

#define DSRT_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) {                                         \
            snprintf(err, err_len, "%s: %s", #call, cudaGetErrorString(e__)); \
            goto fail;                                                    \
        }                                                                 \
    } while (0)

__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

// bounds[0..2] = min of triangle boxes, [3..5] = max, [6..8] = min of centres, [9..11] = max of
// centres, all as ordered uints.
__global__ void k_init_bounds(uint32_t* bounds) {
    int i = threadIdx.x;
    if (i < 12) bounds[i] = ((i / 3) & 1) ? 0u : 0xFFFFFFFFu;
}

__global__ void k_tri_bounds(const DsrtTriangle* __restrict__ tris, int n,
                             float4* __restrict__ lo, float4* __restrict__ hi,
                             uint32_t* __restrict__ bounds) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, h[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float c[3] = {0, 0, 0};
    bool valid = i < n;
    if (valid) {
        const float* p = reinterpret_cast<const float*>(tris + i);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v0 = p[a], v1 = p[3 + a], v2 = p[6 + a];
            l[a] = fminf(fminf(v0, v1), v2);
            h[a] = fmaxf(fmaxf(v0, v1), v2);
            c[a] = 0.5f * l[a] + 0.5f * h[a];
        }
        lo[i] = make_float4(l[0], l[1], l[2], 0.f);
        hi[i] = make_float4(h[0], h[1], h[2], 0.f);
    }
    // warp reduce, then one atomic per warp per component
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float vl = valid ? l[a] : FLT_MAX, vh = valid ? h[a] : -FLT_MAX;
        float cl = valid ? c[a] : FLT_MAX, ch = valid ? c[a] : -FLT_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vl = fminf(vl, __shfl_xor_sync(0xffffffffu, vl, o));
            vh = fmaxf(vh, __shfl_xor_sync(0xffffffffu, vh, o));
            cl = fminf(cl, __shfl_xor_sync(0xffffffffu, cl, o));
            ch = fmaxf(ch, __shfl_xor_sync(0xffffffffu, ch, o));
        }
        if ((threadIdx.x & 31) == 0 && vl <= vh) {
            atomicMin(&bounds[a], float_to_ordered(vl));
            atomicMax(&bounds[3 + a], float_to_ordered(vh));
            atomicMin(&bounds[6 + a], float_to_ordered(cl));
            atomicMax(&bounds[9 + a], float_to_ordered(ch));
        }
    }
}

__device__ __forceinline__ uint64_t spread21(uint32_t v) {
    uint64_t x = v & 0x1FFFFFu;
    x = (x | (x << 32)) & 0x1F00000000FFFFull;
    x = (x | (x << 16)) & 0x1F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, int n,
                         const uint32_t* __restrict__ bounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 l = lo[i], h = hi[i];
    float c[3] = {0.5f * l.x + 0.5f * h.x, 0.5f * l.y + 0.5f * h.y, 0.5f * l.z + 0.5f * h.z};
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float cmin = ordered_to_float(bounds[6 + a]), cmax = ordered_to_float(bounds[9 + a]);
        float ext = cmax - cmin;
        float f = ext > 0.f ? (c[a] - cmin) / ext : 0.f;
        f = fminf(fmaxf(f * 2097152.0f, 0.0f), 2097151.0f);
        q[a] = (uint32_t)f;
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = (uint32_t)i;
}

// ---- binary radix tree (Karras 2012).  Node ids: inner 0..n-2, leaf k -> (n-1)+k. ----------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n,
                            int2* __restrict__ children, int* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? (n - 1 + gamma) : gamma;
    int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

// ---- PLOC: agglomerative clustering along the Morton order ------------------------------------
// Node ids as for the radix tree: leaf k (k-th triangle of the sorted order) is n-1+k; inner nodes
// are handed out downwards from n-2, so that the last merge -- the root -- is node 0.
// Ties in the merged area are broken towards the lower position, which guarantees that the
// globally cheapest pair is mutual (so every round merges at least one pair) and makes the tree,
// and with it every frame, reproducible from run to run.
constexpr int kPlocRadius = 16;          // default search radius
constexpr int kPlocMaxRadius = 64;

__device__ __forceinline__ float union_half_area(float4 l0, float4 h0, float4 l1, float4 h1) {
    const float dx = fmaxf(h0.x, h1.x) - fminf(l0.x, l1.x);
    const float dy = fmaxf(h0.y, h1.y) - fminf(l0.y, l1.y);
    const float dz = fmaxf(h0.z, h1.z) - fminf(l0.z, l1.z);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_ploc_leaves(const uint32_t* __restrict__ vals, const float4* __restrict__ tlo,
                              const float4* __restrict__ thi, int n,
                              float4* __restrict__ blo, float4* __restrict__ bhi, int* __restrict__ clusters) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t tri = vals[k];
    blo[n - 1 + k] = tlo[tri];
    bhi[n - 1 + k] = thi[tri];
    clusters[k] = n - 1 + k;
}

constexpr int kPlocThreads = 256;
__global__ void __launch_bounds__(kPlocThreads)
k_ploc_nearest(const int* __restrict__ clusters, int m, const float4* __restrict__ blo,
               const float4* __restrict__ bhi, int* __restrict__ nearest, int radius) {
    // the boxes of this block's window [base - R, base + T + R) staged in shared memory
    __shared__ float4 s_lo[kPlocThreads + 2 * kPlocMaxRadius];
    __shared__ float4 s_hi[kPlocThreads + 2 * kPlocMaxRadius];
    const int base = blockIdx.x * kPlocThreads;
    for (int t = threadIdx.x; t < kPlocThreads + 2 * radius; t += kPlocThreads) {
        const int j = base - radius + t;
        if (j >= 0 && j < m) {
            const int c = clusters[j];
            s_lo[t] = __ldg(&blo[c]);
            s_hi[t] = __ldg(&bhi[c]);
        }
    }
    __syncthreads();
    const int i = base + threadIdx.x;
    if (i >= m) return;
    const float4 l = s_lo[threadIdx.x + radius], h = s_hi[threadIdx.x + radius];
    float best = FLT_MAX;
    int bj = -1;
    const int j0 = max(0, i - radius), j1 = min(m - 1, i + radius);
    for (int j = j0; j <= j1; ++j) {
        if (j == i) continue;
        const int t = j - base + radius;
        const float a = union_half_area(l, h, s_lo[t], s_hi[t]);
        if (a < best) { best = a; bj = j; }        // ascending j and a strict test: lowest position wins ties
    }
    nearest[i] = bj;
}

// flags[i]: low word 1 = position i survives the round (it is not the upper half of a merging
// pair), high word 1 = position i is the lower half of a merging pair and creates the new node.
__global__ void k_ploc_mark(const int* __restrict__ nearest, int m, unsigned long long* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nearest[i];
    const bool mutual = j >= 0 && nearest[j] == i;
    const unsigned long long keep = (mutual && j < i) ? 0ull : 1ull;
    const unsigned long long lead = (mutual && i < j) ? 1ull : 0ull;
    flags[i] = keep | (lead << 32);
}

// counters[0] = clusters left, counters[1] = next inner node id (counts down).
__global__ void k_ploc_merge(const int* __restrict__ clusters, const int* __restrict__ nearest, int m,
                             const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ scan,
                             int next_id, int2* __restrict__ children, int* __restrict__ parent,
                             float4* __restrict__ blo, float4* __restrict__ bhi,
                             int* __restrict__ clusters_out, int* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long f = flags[i], sc = scan[i];
    const int pos = (int)(sc & 0xFFFFFFFFull), rank = (int)(sc >> 32);
    if (i == m - 1) {
        counters[0] = pos + (int)(f & 1ull);
        counters[1] = next_id - (rank + (int)(f >> 32));
    }
    if ((f & 1ull) == 0ull) return;
    const int c = clusters[i];
    if (f >> 32) {
        const int d = clusters[nearest[i]];
        const int id = next_id - rank;
        children[id] = make_int2(c, d);
        parent[c] = id;
        parent[d] = id;
        const float4 l0 = blo[c], h0 = bhi[c], l1 = blo[d], h1 = bhi[d];
        blo[id] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        bhi[id] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        clusters_out[pos] = id;
    } else {
        clusters_out[pos] = c;
    }
}

// ---- bottom-up refit + optimal wide-node assignment ------------------------------------------
// Besides the boxes, each binary node gets the surface-area cost of representing its subtree
// with at most i = 1..7 child slots of a wide node, and the decisions that achieve it
// (dynamic programme of Ylitie et al. 2017, section 4.1, with c_node = 1, c_prim = 0.3):
//   C(n,1)   = min( A_n * P_n * c_prim  if P_n <= 3          [one leaf child],
//                   A_n * c_node + D(n,8)                     [a wide node of its own] )
//   C(n,i)   = min( D(n,i), C(n,i-1) )                        i = 2..7
//   D(n,j)   = min over 0<k<j of C(left,k) + C(right,j-k)
// dec packs 3 bits per i: i=1: 0 leaf / 1 wide node; i=2..7: 0 = "use i-1", else k; i=8: k.
constexpr float kCostNode = 1.0f;
constexpr float kCostPrimDefault = 0.3f;

__device__ __forceinline__ float box_half_area(float4 l, float4 h) {
    float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_refit(const uint32_t* __restrict__ vals, const float4* __restrict__ tlo,
                        const float4* __restrict__ thi, int n,
                        const int2* __restrict__ children, const int* __restrict__ parent,
                        int* __restrict__ count,
                        float4* __restrict__ blo, float4* __restrict__ bhi, int* __restrict__ flags,
                        float* __restrict__ cost, uint32_t* __restrict__ dec, float kCostPrim) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t tri = vals[k];
    float4 l = tlo[tri], h = thi[tri];
    int node = n - 1 + k;
    blo[node] = l;
    bhi[node] = h;
    count[node] = 1;
    {
        const float c = box_half_area(l, h) * kCostPrim;
        float* cn = cost + (size_t)node * 8;
        for (int i = 1; i <= 7; ++i) cn[i] = c;
        dec[node] = 0u;                              // a single triangle is always a leaf child
    }
    if (n == 1) return;
    int p = parent[node];
    while (p >= 0) {
        __threadfence();
        if (atomicAdd(&flags[p], 1) == 0) return;     // the sibling subtree finishes this node
        int2 ch = children[p];
        float4 l0 = __ldcg(&blo[ch.x]), h0 = __ldcg(&bhi[ch.x]);
        float4 l1 = __ldcg(&blo[ch.y]), h1 = __ldcg(&bhi[ch.y]);
        l = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        h = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        blo[p] = l;
        bhi[p] = h;
        // dynamic programme
        float cl[8], cr[8];
        for (int i = 1; i <= 7; ++i) {
            cl[i] = __ldcg(&cost[(size_t)ch.x * 8 + i]);
            cr[i] = __ldcg(&cost[(size_t)ch.y * 8 + i]);
        }
        const float area = box_half_area(l, h);
        const int cnt = __ldcg(&count[ch.x]) + __ldcg(&count[ch.y]);
        count[p] = cnt;
        float dist[9];
        uint32_t kbest[9];
        for (int j = 2; j <= 8; ++j) {
            float best = FLT_MAX;
            uint32_t bk = 1;
            for (int kk = 1; kk < j; ++kk) {
                if (kk > 7 || j - kk > 7) continue;
                float c = cl[kk] + cr[j - kk];
                if (c < best) { best = c; bk = (uint32_t)kk; }
            }
            dist[j] = best;
            kbest[j] = bk;
        }
        uint32_t d = kbest[8] << 24;
        float cp[8];
        const float c_leaf = cnt <= kMaxLeafTris ? area * (float)cnt * kCostPrim : FLT_MAX;
        const float c_wide = area * kCostNode + dist[8];
        if (c_leaf <= c_wide) { cp[1] = c_leaf; } else { cp[1] = c_wide; d |= 1u << 3; }
        for (int i = 2; i <= 7; ++i) {
            if (dist[i] < cp[i - 1]) { cp[i] = dist[i]; d |= kbest[i] << (3 * i); }
            else cp[i] = cp[i - 1];
        }
        float* cn = cost + (size_t)p * 8;
        for (int i = 1; i <= 7; ++i) cn[i] = cp[i];
        dec[p] = d;
        p = parent[p];
    }
}

// ---- collapse -------------------------------------------------------------------------------
struct CollapseArgs {
    int n;                         // triangles
    const int2* children;
    const int* count;              // triangles below each binary node
    const float4* blo;
    const float4* bhi;
    const uint32_t* dec;           // per binary node: decisions of the dynamic programme (k_refit)
    const uint32_t* vals;          // sorted position -> original triangle index
    const DsrtTriangle* tris;      // raw triangles (device copy)
    uint4* nodes;                  // out
    float4* tri_records;           // out
    uint32_t* counters;            // [0] nodes allocated, [1] triangle records allocated,
                                   // [2] next-level item count, [3] float bits of SAH accumulator
    float eps;                     // absolute box padding
    float inv_root_area;
};

__device__ __forceinline__ float half_area(float4 l, float4 h) {
    float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ int subtree_count(const CollapseArgs& A, int node) {
    return node >= A.n - 1 ? 1 : A.count[node];
}

// The (at most kMaxLeafTris) sorted positions below `node`, left to right.
__device__ __forceinline__ int subtree_leaves(const CollapseArgs& A, int node, int* out) {
    int st[kMaxLeafTris + 1];
    int sp = 0, nl = 0;
    st[sp++] = node;
    while (sp > 0) {
        const int c = st[--sp];
        if (c >= A.n - 1) { if (nl < kMaxLeafTris) out[nl] = c - (A.n - 1); ++nl; continue; }
        const int2 cc = A.children[c];
        st[sp++] = cc.y;
        st[sp++] = cc.x;
    }
    return nl;
}

__global__ void k_collapse_level(CollapseArgs A, const int2* __restrict__ items_in, int count_in,
                                 int2* __restrict__ items_out) {
    int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= count_in) return;
    const int bnode = items_in[it].x;
    const int wnode = items_in[it].y;
    const int n = A.n;

    // Children of this wide node: follow the decisions of the dynamic programme.  `leafbits`
    // marks the children that are leaf clusters; the others become wide nodes themselves.
    int ch[8];
    int nch = 0;
    uint32_t leafbits = 0;
    {
        const bool root_is_leaf = bnode >= n - 1 || ((A.dec[bnode] >> 3) & 7u) == 0u;
        if (root_is_leaf) {                       // the whole scene fits one leaf child (n <= 3)
            ch[0] = bnode; nch = 1; leafbits = 1u;
        } else {
            int st_node[16], st_slots[16];
            int sp = 0;
            st_node[sp] = bnode; st_slots[sp] = 8; ++sp;          // distribute bnode's children over 8 slots
            bool first = true;
            while (sp > 0) {
                --sp;
                int c = st_node[sp], i = st_slots[sp];
                if (!first) {
                    // represent subtree c with at most i slots
                    const uint32_t d = c >= n - 1 ? 0u : A.dec[c];
                    while (i > 1 && ((d >> (3 * i)) & 7u) == 0u) --i;
                    if (i == 1 || c >= n - 1) {
                        const bool leaf = c >= n - 1 || ((d >> 3) & 7u) == 0u;
                        if (leaf) leafbits |= 1u << nch;
                        ch[nch++] = c;
                        continue;
                    }
                }
                first = false;
                const uint32_t d = A.dec[c];
                const int k = (int)((d >> (3 * i)) & 7u);
                const int2 cc = A.children[c];
                st_node[sp] = cc.y; st_slots[sp] = i - k; ++sp;
                st_node[sp] = cc.x; st_slots[sp] = k; ++sp;
            }
        }
    }

    const float4 nlo = A.blo[bnode], nhi = A.bhi[bnode];
    const float eps = A.eps;
    const float px = nlo.x - eps, py = nlo.y - eps, pz = nlo.z - eps;
    const float cx = 0.5f * (nlo.x + nhi.x), cy = 0.5f * (nlo.y + nhi.y), cz = 0.5f * (nlo.z + nhi.z);

    // ---- octant-matched slot assignment (greedy on the 8x8 cost table)
    float4 clo[8], chi[8];
    float ox[8], oy[8], oz[8];
    for (int k = 0; k < nch; ++k) {
        clo[k] = A.blo[ch[k]]; chi[k] = A.bhi[ch[k]];
        ox[k] = 0.5f * (clo[k].x + chi[k].x) - cx;
        oy[k] = 0.5f * (clo[k].y + chi[k].y) - cy;
        oz[k] = 0.5f * (clo[k].z + chi[k].z) - cz;
    }
    int slot_child[8];
    for (int s = 0; s < 8; ++s) slot_child[s] = -1;
    uint32_t child_done = 0, slot_done = 0;
    for (int round = 0; round < nch; ++round) {
        float best = -FLT_MAX;
        int bc = -1, bs = -1;
        for (int k = 0; k < nch; ++k) {
            if (child_done & (1u << k)) continue;
            for (int s = 0; s < 8; ++s) {
                if (slot_done & (1u << s)) continue;
                float cost = ((s & 1) ? ox[k] : -ox[k]) + ((s & 2) ? oy[k] : -oy[k]) + ((s & 4) ? oz[k] : -oz[k]);
                if (cost > best) { best = cost; bc = k; bs = s; }
            }
        }
        slot_child[bs] = bc;
        child_done |= 1u << bc;
        slot_done |= 1u << bs;
    }

    // ---- classify, count, allocate
    uint32_t imask = 0;
    int n_inner = 0, n_tris = 0;
    for (int s = 0; s < 8; ++s) {
        int k = slot_child[s];
        if (k < 0) continue;
        if ((leafbits >> k) & 1u) n_tris += subtree_count(A, ch[k]);
        else { imask |= 1u << s; ++n_inner; }
    }
    uint32_t child_base = 0, tri_base = 0, item_base = 0;
    if (n_inner) {
        child_base = atomicAdd(&A.counters[0], (uint32_t)n_inner);
        item_base = atomicAdd(&A.counters[2], (uint32_t)n_inner);
    }
    if (n_tris) tri_base = atomicAdd(&A.counters[1], (uint32_t)n_tris);

    // ---- quantisation grid
    int e[3];
    float scale[3];
    {
        float ext[3] = {(nhi.x + eps) - px, (nhi.y + eps) - py, (nhi.z + eps) - pz};
        for (int a = 0; a < 3; ++a) {
            int ee;
            frexpf(ext[a] / 255.0f, &ee);          // 2^ee > ext/255
            if (ee < -125) ee = -125;
            while (ldexpf(255.0f, ee) < ext[a]) ++ee;
            e[a] = ee;
            scale[a] = ldexpf(1.0f, ee);
        }
    }
    const float p3[3] = {px, py, pz};

    uint32_t meta[8];
    uint32_t weight[8];        // bfloat16 bits of unary(count) << offset for leaf children, else 0 (bvh8.cuh: w5)
    uint32_t qlo[3][8], qhi[3][8];
    int rank = 0, tri_off = 0;
    float sah = half_area(nlo, nhi) * (float)nch * 0.5f;    // cost of testing this node's boxes
    for (int s = 0; s < 8; ++s) {
        int k = slot_child[s];
        if (k < 0) {
            meta[s] = 0;
            weight[s] = 0;
            for (int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; }
            continue;
        }
        const float l3[3] = {clo[k].x, clo[k].y, clo[k].z};
        const float h3[3] = {chi[k].x, chi[k].y, chi[k].z};
        for (int a = 0; a < 3; ++a) {
            float fl = floorf((l3[a] - eps - p3[a]) / scale[a]);
            int ql = (int)fminf(fmaxf(fl, 0.f), 255.f);
            while (ql > 0 && fmaf((float)ql, scale[a], p3[a]) > l3[a] - 0.5f * eps) --ql;
            float fh = ceilf((h3[a] + eps - p3[a]) / scale[a]);
            int qh = (int)fminf(fmaxf(fh, 0.f), 255.f);
            while (qh < 255 && fmaf((float)qh, scale[a], p3[a]) < h3[a] + 0.5f * eps) ++qh;
            qlo[a][s] = (uint32_t)ql;
            qhi[a][s] = (uint32_t)qh;
        }
        int node_c = ch[k];
        int cnt = subtree_count(A, node_c);
        if ((leafbits >> k) & 1u) {
            int leaf_pos[kMaxLeafTris];
            subtree_leaves(A, node_c, leaf_pos);
            for (int j = 0; j < cnt; ++j) {
                uint32_t tri = A.vals[leaf_pos[j]];
                const float* tp = reinterpret_cast<const float*>(A.tris + tri);
                float v0x = tp[0], v0y = tp[1], v0z = tp[2];
                float4* rec = A.tri_records + (size_t)(tri_base + tri_off + j) * kTriWords;
                rec[0] = make_float4(v0x, v0y, v0z, __uint_as_float(tri));
                rec[1] = make_float4(__fsub_rn(tp[3], v0x), __fsub_rn(tp[4], v0y), __fsub_rn(tp[5], v0z), 0.f);
                rec[2] = make_float4(__fsub_rn(tp[6], v0x), __fsub_rn(tp[7], v0y), __fsub_rn(tp[8], v0z), 0.f);
            }
            meta[s] = (((1u << cnt) - 1u) << 5) | (uint32_t)tri_off;
            // (2^cnt - 1) * 2^tri_off = 1.0, 1.5 or 1.75 times 2^(tri_off + cnt - 1): the top 16 bits of that float
            weight[s] = ((uint32_t)(127 + tri_off + cnt - 1) << 7) | (cnt == 1 ? 0x00u : cnt == 2 ? 0x40u : 0x60u);
            tri_off += cnt;
            sah += half_area(clo[k], chi[k]) * (float)cnt;
        } else {
            meta[s] = 0x20u | (24u + (uint32_t)s);
            weight[s] = 0;
            items_out[item_base + rank] = make_int2(node_c, (int)(child_base + rank));
            ++rank;
        }
    }
    atomicAdd(reinterpret_cast<float*>(&A.counters[3]), sah * A.inv_root_area);

    auto pack4 = [](const uint32_t* v) { return v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24); };
    uint4* np = A.nodes + (size_t)wnode * kNodeWords;
    np[0] = make_uint4(__float_as_uint(px), __float_as_uint(py), __float_as_uint(pz),
                       (uint32_t)(e[0] + 127) | ((uint32_t)(e[1] + 127) << 8) |
                           ((uint32_t)(e[2] + 127) << 16) | (imask << 24));
    np[1] = make_uint4(child_base, tri_base, pack4(meta), pack4(meta + 4));
    np[2] = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    np[3] = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    np[4] = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    np[5] = make_uint4(weight[0] | (weight[1] << 16), weight[2] | (weight[3] << 16),
                       weight[4] | (weight[5] << 16), weight[6] | (weight[7] << 16));
}

// ---- per-triangle shading records (indexed by the uploaded triangle index) ---------------
//   s0: n0.xyz, bits(material_id)   s1: n1.xyz, bits(albedo_tex)   s2: n2.xyz, 0
//   uv (only when textures exist): uv0.xy, uv1.xy | uv2.xy, 0, 0
__global__ void k_shade_records(const DsrtTriangle* __restrict__ tris, int n,
                                float4* __restrict__ shade, float4* __restrict__ uvrec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(tris + i);
    const int* pi = reinterpret_cast<const int*>(tris + i);
    shade[(size_t)i * 3 + 0] = make_float4(p[9], p[10], p[11], __int_as_float(pi[27]));
    shade[(size_t)i * 3 + 1] = make_float4(p[12], p[13], p[14], __int_as_float(pi[28]));
    shade[(size_t)i * 3 + 2] = make_float4(p[15], p[16], p[17], 0.f);
    if (uvrec) {
        uvrec[(size_t)i * 2 + 0] = make_float4(p[18], p[19], p[21], p[22]);
        uvrec[(size_t)i * 2 + 1] = make_float4(p[24], p[25], 0.f, 0.f);
    }
}

// flag[0] is set when some triangle cannot use the one-word shading record: its three vertex normals
// differ in some bit, or an id does not fit 16 bits.
__global__ void k_check_flat_normals(const DsrtTriangle* __restrict__ tris, int n, uint32_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(tris + i);
    const int* pi = reinterpret_cast<const int*>(tris + i);
    bool bad = false;
    for (int k = 0; k < 3; ++k) bad = bad || p[9 + k] != p[12 + k] || p[9 + k] != p[15 + k];
    bad = bad || pi[27] < 0 || pi[27] >= 65535 || pi[28] < -1 || pi[28] >= 65534;
    if (bad) flag[0] = 1u;
}

__global__ void k_shade_records_compact(const DsrtTriangle* __restrict__ tris, int n,
                                        float4* __restrict__ shade, float4* __restrict__ uvrec) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = reinterpret_cast<const float*>(tris + i);
    const int* pi = reinterpret_cast<const int*>(tris + i);
    shade[i] = make_float4(p[9], p[10], p[11], __uint_as_float((uint32_t)pi[27] | ((uint32_t)(pi[28] + 1) << 16)));
    if (uvrec) {
        uvrec[(size_t)i * 2 + 0] = make_float4(p[18], p[19], p[21], p[22]);
        uvrec[(size_t)i * 2 + 1] = make_float4(p[24], p[25], 0.f, 0.f);
    }
}

}  // namespace

void free_bvh8(Bvh8& b) {
    if (b.nodes) cudaFree(b.nodes);
    if (b.tri_records) cudaFree(b.tri_records);
    if (b.shade_records) cudaFree(b.shade_records);
    if (b.uv_records) cudaFree(b.uv_records);
    b = Bvh8();
}

int build_bvh8(cudaStream_t stream, const DsrtTriangle* d_tris, int n, bool want_uv, int builder,
               Bvh8* out, char* err, size_t err_len) {
    Bvh8 B;
    float4 *tlo = nullptr, *thi = nullptr, *blo = nullptr, *bhi = nullptr;
    uint32_t *bounds = nullptr, *vals = nullptr, *vals_s = nullptr, *counters = nullptr;
    uint64_t *keys = nullptr, *keys_s = nullptr;
    int2 *children = nullptr, *items[2] = {nullptr, nullptr};
    int *parent = nullptr, *flags = nullptr, *count = nullptr;
    int *clusters[2] = {nullptr, nullptr}, *nearest = nullptr, *ploc_counters = nullptr;
    unsigned long long *pflags = nullptr, *pscan = nullptr;
    void* scan_tmp = nullptr;
    size_t scan_bytes = 0;
    int ploc_rounds = 0;
    float* cost = nullptr;
    uint32_t* dec = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    const int TB = 256;
    const int nb = (n + TB - 1) / TB;
    uint32_t hb[12];
    uint32_t hc[4];
    int level_count = 1, levels = 0;
    float cost_prim = kCostPrimDefault;
    if (const char* e = getenv("DSRT_COST_PRIM")) cost_prim = (float)atof(e);       // tuning experiments only
    size_t max_nodes = 0;

    if (n <= 0) { *out = B; return 0; }

    DSRT_CUDA(cudaEventCreate(&ev0));
    DSRT_CUDA(cudaEventCreate(&ev1));
    DSRT_CUDA(cudaEventRecord(ev0, stream));

    DSRT_CUDA(cudaMalloc(&tlo, sizeof(float4) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&thi, sizeof(float4) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&bounds, sizeof(uint32_t) * 12));
    DSRT_CUDA(cudaMalloc(&keys, sizeof(uint64_t) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&keys_s, sizeof(uint64_t) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&vals, sizeof(uint32_t) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&vals_s, sizeof(uint32_t) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&children, sizeof(int2) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&count, sizeof(int) * (size_t)(2 * n)));
    DSRT_CUDA(cudaMalloc(&parent, sizeof(int) * (size_t)(2 * n)));
    DSRT_CUDA(cudaMalloc(&flags, sizeof(int) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&blo, sizeof(float4) * (size_t)(2 * n)));
    DSRT_CUDA(cudaMalloc(&bhi, sizeof(float4) * (size_t)(2 * n)));
    DSRT_CUDA(cudaMalloc(&cost, sizeof(float) * 8 * (size_t)(2 * n)));
    DSRT_CUDA(cudaMalloc(&dec, sizeof(uint32_t) * (size_t)(2 * n)));
    DSRT_CUDA(cudaMalloc(&counters, sizeof(uint32_t) * 4));
    DSRT_CUDA(cudaMalloc(&items[0], sizeof(int2) * (size_t)n));
    DSRT_CUDA(cudaMalloc(&items[1], sizeof(int2) * (size_t)n));

    k_init_bounds<<<1, 32, 0, stream>>>(bounds);
    k_tri_bounds<<<nb, TB, 0, stream>>>(d_tris, n, tlo, thi, bounds);
    k_morton<<<nb, TB, 0, stream>>>(tlo, thi, n, bounds, keys, vals);
    DSRT_CUDA(cudaGetLastError());

    DSRT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys, keys_s, vals, vals_s, n, 0, 63, stream));
    DSRT_CUDA(cudaMalloc(&cub_tmp, cub_bytes ? cub_bytes : 16));
    DSRT_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys, keys_s, vals, vals_s, n, 0, 63, stream));

    DSRT_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)n, stream));
    if (n > 1 && builder == DSRT_BUILDER_LBVH) {
        k_hierarchy<<<(n - 1 + TB - 1) / TB, TB, 0, stream>>>(keys_s, n, children, parent);
    } else if (n > 1) {
        DSRT_CUDA(cudaMalloc(&clusters[0], sizeof(int) * (size_t)n));
        DSRT_CUDA(cudaMalloc(&clusters[1], sizeof(int) * (size_t)n));
        DSRT_CUDA(cudaMalloc(&nearest, sizeof(int) * (size_t)n));
        DSRT_CUDA(cudaMalloc(&pflags, sizeof(unsigned long long) * (size_t)n));
        DSRT_CUDA(cudaMalloc(&pscan, sizeof(unsigned long long) * (size_t)n));
        DSRT_CUDA(cudaMalloc(&ploc_counters, sizeof(int) * 2));
        DSRT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, pflags, pscan, n, stream));
        DSRT_CUDA(cudaMalloc(&scan_tmp, scan_bytes ? scan_bytes : 16));
        k_ploc_leaves<<<nb, TB, 0, stream>>>(vals_s, tlo, thi, n, blo, bhi, clusters[0]);
        int m = n, next_id = n - 2, cur = 0;
        int radius = kPlocRadius;
        if (const char* e = getenv("DSRT_PLOC_RADIUS")) radius = max(1, min(kPlocMaxRadius, atoi(e)));     // tuning experiments only
        while (m > 1) {
            const int mb = (m + kPlocThreads - 1) / kPlocThreads;
            k_ploc_nearest<<<mb, kPlocThreads, 0, stream>>>(clusters[cur], m, blo, bhi, nearest, radius);
            k_ploc_mark<<<mb, kPlocThreads, 0, stream>>>(nearest, m, pflags);
            DSRT_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, pflags, pscan, m, stream));
            k_ploc_merge<<<mb, kPlocThreads, 0, stream>>>(clusters[cur], nearest, m, pflags, pscan, next_id, children, parent,
                                                          blo, bhi, clusters[cur ^ 1], ploc_counters);
            DSRT_CUDA(cudaGetLastError());
            int hc2[2];
            DSRT_CUDA(cudaMemcpyAsync(hc2, ploc_counters, sizeof(hc2), cudaMemcpyDeviceToHost, stream));
            DSRT_CUDA(cudaStreamSynchronize(stream));
            if (hc2[0] >= m) { snprintf(err, err_len, "PLOC made no progress at %d clusters", m); goto fail; }
            m = hc2[0];
            next_id = hc2[1];
            cur ^= 1;
            ++ploc_rounds;
        }
        if (next_id != -1) { snprintf(err, err_len, "PLOC created %d inner nodes for %d triangles", n - 2 - next_id, n); goto fail; }
        DSRT_CUDA(cudaMemsetAsync(parent, 0xFF, sizeof(int), stream));       // the root (node 0) has parent -1
    }
    k_refit<<<nb, TB, 0, stream>>>(vals_s, tlo, thi, n, children, parent, count, blo, bhi, flags, cost, dec, cost_prim);
    DSRT_CUDA(cudaGetLastError());

    DSRT_CUDA(cudaMemcpyAsync(hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    DSRT_CUDA(cudaStreamSynchronize(stream));

    // Output arrays.  A wide node has >= 2 children unless it is the lone root, and every
    // inner child of a wide node holds >= 4 triangles, so n/2 + 1 nodes always suffice.
    max_nodes = (size_t)n / 2 + 2;
    DSRT_CUDA(cudaMalloc(&B.nodes, sizeof(uint4) * kNodeWords * max_nodes));
    DSRT_CUDA(cudaMalloc(&B.tri_records, sizeof(float4) * kTriWords * (size_t)n));
    {
        // one-word shading records when every triangle has a single normal (always, with the reference's loader)
        uint32_t nonflat = 0u;
        DSRT_CUDA(cudaMemsetAsync(counters, 0, sizeof(uint32_t), stream));
        k_check_flat_normals<<<nb, TB, 0, stream>>>(d_tris, n, counters);
        DSRT_CUDA(cudaMemcpyAsync(&nonflat, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        DSRT_CUDA(cudaStreamSynchronize(stream));
        B.shade_compact = nonflat == 0u && getenv("DSRT_FULL_SHADE_RECORDS") == nullptr;
    }
    DSRT_CUDA(cudaMalloc(&B.shade_records, sizeof(float4) * (B.shade_compact ? 1 : 3) * (size_t)n));
    if (want_uv) DSRT_CUDA(cudaMalloc(&B.uv_records, sizeof(float4) * 2 * (size_t)n));
    if (B.shade_compact) k_shade_records_compact<<<nb, TB, 0, stream>>>(d_tris, n, B.shade_records, B.uv_records);
    else k_shade_records<<<nb, TB, 0, stream>>>(d_tris, n, B.shade_records, B.uv_records);

    {
        float ext = 0.f;
        for (int a = 0; a < 3; ++a) {
            B.scene_lo[a] = ordered_to_float(hb[a]);
            B.scene_hi[a] = ordered_to_float(hb[3 + a]);
            ext = fmaxf(ext, B.scene_hi[a] - B.scene_lo[a]);
            ext = fmaxf(ext, fmaxf(fabsf(B.scene_lo[a]), fabsf(B.scene_hi[a])));
        }
        if (!(ext > 0.f)) ext = 1.f;
        CollapseArgs A;
        A.n = n; A.children = children; A.count = count; A.blo = blo; A.bhi = bhi;
        A.dec = dec; A.vals = vals_s; A.tris = d_tris; A.nodes = B.nodes; A.tri_records = B.tri_records;
        A.counters = counters;
        A.eps = ext * (1.0f / 524288.0f);         // 2^-19 of the scene size
        {
            float dx = B.scene_hi[0] - B.scene_lo[0], dy = B.scene_hi[1] - B.scene_lo[1], dz = B.scene_hi[2] - B.scene_lo[2];
            float ra = dx * dy + dy * dz + dz * dx;
            A.inv_root_area = ra > 0.f ? 1.0f / ra : 0.f;
        }
        hc[0] = 1; hc[1] = 0; hc[2] = 0; hc[3] = 0;
        DSRT_CUDA(cudaMemcpyAsync(counters, hc, sizeof(hc), cudaMemcpyHostToDevice, stream));
        // binary root: inner node 0, or the single leaf when n == 1
        int2 root_item = make_int2(n == 1 ? 0 : 0, 0);
        if (n == 1) root_item.x = 0;               // leaf id n-1+0 == 0
        DSRT_CUDA(cudaMemcpyAsync(items[0], &root_item, sizeof(int2), cudaMemcpyHostToDevice, stream));
        int cur = 0;
        while (level_count > 0) {
            ++levels;
            k_collapse_level<<<(level_count + 127) / 128, 128, 0, stream>>>(A, items[cur], level_count, items[cur ^ 1]);
            DSRT_CUDA(cudaGetLastError());
            DSRT_CUDA(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, stream));
            DSRT_CUDA(cudaStreamSynchronize(stream));
            level_count = (int)hc[2];
            uint32_t zero = 0;
            DSRT_CUDA(cudaMemcpyAsync(counters + 2, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
            cur ^= 1;
            if (levels > kSmemStack + kLocalStack) {
                snprintf(err, err_len, "wide BVH deeper than the traversal stack (%d levels)", levels);
                goto fail;
            }
        }
        B.num_nodes = (int)hc[0];
        B.num_tris = n;
        B.num_levels = levels;
        B.build_rounds = ploc_rounds;
        memcpy(&B.sah_cost, &hc[3], 4);
        if ((int)hc[1] != n) {
            snprintf(err, err_len, "collapse emitted %u triangle records for %d triangles", hc[1], n);
            goto fail;
        }
    }

    DSRT_CUDA(cudaEventRecord(ev1, stream));
    DSRT_CUDA(cudaEventSynchronize(ev1));
    DSRT_CUDA(cudaEventElapsedTime(&B.build_ms, ev0, ev1));

    cudaFree(tlo); cudaFree(thi); cudaFree(bounds); cudaFree(keys); cudaFree(keys_s);
    cudaFree(vals); cudaFree(vals_s); cudaFree(children); cudaFree(count); cudaFree(parent);
    cudaFree(clusters[0]); cudaFree(clusters[1]); cudaFree(nearest); cudaFree(pflags); cudaFree(pscan);
    cudaFree(ploc_counters); cudaFree(scan_tmp);
    cudaFree(flags); cudaFree(blo); cudaFree(bhi); cudaFree(counters); cudaFree(cost); cudaFree(dec);
    cudaFree(items[0]); cudaFree(items[1]); cudaFree(cub_tmp);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    *out = B;
    return 0;

fail:
    cudaFree(tlo); cudaFree(thi); cudaFree(bounds); cudaFree(keys); cudaFree(keys_s);
    cudaFree(vals); cudaFree(vals_s); cudaFree(children); cudaFree(count); cudaFree(parent);
    cudaFree(clusters[0]); cudaFree(clusters[1]); cudaFree(nearest); cudaFree(pflags); cudaFree(pscan);
    cudaFree(ploc_counters); cudaFree(scan_tmp);
    cudaFree(flags); cudaFree(blo); cudaFree(bhi); cudaFree(counters); cudaFree(cost); cudaFree(dec);
    cudaFree(items[0]); cudaFree(items[1]); cudaFree(cub_tmp);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    free_bvh8(B);
    return -1;
}

}  // namespace dsrt
