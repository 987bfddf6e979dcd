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

__device__ __forceinline__ float signf(float v) { return (float)((v > 0.f) - (v < 0.f)); }

}  // namespace vl3d
