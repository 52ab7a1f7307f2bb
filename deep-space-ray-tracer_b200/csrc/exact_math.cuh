// exact_math.cuh -- float3 arithmetic with the rounding spelled out.
//
// "Bit-exact first hit" against the reference means reproducing how nvcc 12.9 contracts the
// reference's float expressions for sm_100a (re-derived from
// `cuobjdump -sass` of /root/reference/src/gpu_render.cu built with -rdc, default -fmad):
//
//   a.x*b.x + a.y*b.y + a.z*b.z     ->  fma(a.z,b.z, fma(a.x,b.x, rn(a.y*b.y)))
//   p*q - r*s   (cross components)  ->  fma(p,q, -rn(r*s))
//   A*s + B*t                       ->  fma(A,s, rn(B*t))      (first product fused)
//   c + A*s ,  A*s + c              ->  fma(A,s,c)
//   c - A*s                         ->  fma(-A,s,c)
//   1.0f/x, x/y, sqrtf              ->  IEEE round-to-nearest (no fast-math)
//
// Every helper below uses __fmaf_rn/__fmul_rn/__fadd_rn, which the compiler never
// re-associates or re-contracts, so the result does not depend on optimisation context.
#pragma once

#include <cuda_runtime.h>

namespace dsrt {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }

__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return v3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return v3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vmul(V3 a, V3 b) { return v3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vscale(V3 a, float s) { return v3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ V3 vneg(V3 a) { return v3(-a.x, -a.y, -a.z); }

// reference f3_dot (gpu_render.cu:31-33) as compiled
__device__ __forceinline__ float vdot(V3 a, V3 b) {
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}
// reference f3_cross (gpu_render.cu:35-41) as compiled
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
    return v3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)),
              __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)),
              __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
// a + b*s  (ray.at, shadow-origin offset, fuzz perturbation)
__device__ __forceinline__ V3 vmadd(V3 b, float s, V3 a) {
    return v3(__fmaf_rn(b.x, s, a.x), __fmaf_rn(b.y, s, a.y), __fmaf_rn(b.z, s, a.z));
}
// reference f3_norm (gpu_render.cu:51-56): zero vector when the length is <= 0
__device__ __forceinline__ V3 vnorm(V3 a) {
    float L = __fsqrt_rn(vdot(a, a));
    if (!(L > 0.0f)) return v3(0.0f, 0.0f, 0.0f);       // `L <= 0` false for NaN -> falls through below
    float inv = __frcp_rn(L);
    return vscale(a, inv);
}
// The reference's test is `if (L <= 0.0f) return 0`; for NaN lengths it keeps going and
// produces NaNs.  vnorm_ref keeps that behaviour for the bit-replay path.
__device__ __forceinline__ V3 vnorm_ref(V3 a) {
    float L = __fsqrt_rn(vdot(a, a));
    if (L <= 0.0f) return v3(0.0f, 0.0f, 0.0f);
    float inv = __frcp_rn(L);
    return vscale(a, inv);
}
__device__ __forceinline__ float clamp01(float x) { return fminf(1.0f, fmaxf(0.0f, x)); }

}  // namespace dsrt
