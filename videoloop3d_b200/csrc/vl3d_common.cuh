// Shared device/host helpers for libvl3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "vl3d.h"

namespace vl3d {

// thread-local last-error text (vl3d_last_error_string)
char* err_buf();
int set_err(int code, const char* fmt, ...);
int check_launch(const char* what);

#define VL3D_REQUIRE(cond, code, ...)                          \
    do {                                                       \
        if (!(cond)) return vl3d::set_err((code), __VA_ARGS__); \
    } while (0)

__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// vectorised no-return reduction: one 16-byte RED per texel (sm_90+)
__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One Adam element update (MPV.py:200-218, torch.optim.Adam's single-tensor order):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= step_size * m / (sqrt(v) * inv_sqrt_bc2 + eps)
// with MUFU sqrt / reciprocal (each ~1 ulp; the update agrees with the IEEE formulation to ~4e-7 relative, i.e. 4e-9
// absolute at lr = 0.01) instead of the ~45-instruction IEEE sqrt + division sequences: inside the fused backward + Adam
// kernel the optimiser arithmetic competes with the issue-bound backward for issue slots.  v >= 0 and the denominator
// >= eps, so flush-to-zero never changes a result by more than eps * 1e-30.  Every Adam path of the library uses this.
__device__ __forceinline__ void adam1(float& p, const float g, float& m, float& v, const float b1, const float b2,
                                      const float step_size, const float inv_sqrt_bc2, const float eps) {
    m = b1 * m + (1.f - b1) * g;
    v = b2 * v + (1.f - b2) * g * g;
    float s, r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(v));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(s, inv_sqrt_bc2, eps)));
    p -= step_size * (m * r);
}

__device__ __forceinline__ float signf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

}  // namespace vl3d
