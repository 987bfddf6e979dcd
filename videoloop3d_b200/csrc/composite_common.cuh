// Shared pieces of the composite kernels (composite.cu, fused_bwd_adam.cu): tile shape, the kernel parameter block and
// the geometry helpers of the slot-ordered forward.
//
// Mapping: one thread = one screen pixel, 32 lanes = 32 adjacent pixels of a row, so the four
// bilinear taps of a warp are four coalesced 512-byte runs of RGBA texels (one LDG.128 per tap).
// A CTA is a 32x8 pixel tile x a chunk of TF frames; geometry (hit mask, tap address, bilinear
// weights) depends on (pixel, plane) only and is shared by the TF frames in registers.
#pragma once
#include <stdlib.h>

#include "vl3d_common.cuh"

namespace vl3d {

constexpr int BX = 32;
constexpr int BY = 8;

struct CompositeParams {
    vl3d_view view;
    const vl3d_quad* quads;
    const float4* atlas_dyn;
    const float4* atlas_sta;
    const int* ts;
    int T, pad;
    int tb;          // first frame of this launch (lean kernels: a call is split into a TF-multiple + a tail)
    // forward
    float* rgb_out;
    float* alpha_out;
    double* smooth;
    float4* mpi_out;
    int* hits_out;
    // backward
    const float* grad_rgb;
    const float* rgb;
    float4* grad_dyn;
    float4* grad_sta;
    const float* w_smooth;
};

// quad-grid coordinates of pixel (u, v) on plane with homography h; false if behind / outside.
__device__ __forceinline__ bool plane_grid(const float* __restrict__ h, float u, float v, int qw, int qh,
                                           float& gx, float& gy) {
    const float w = fmaf(h[6], u, fmaf(h[7], v, h[8]));
    float inv = __fdividef(1.f, w);                   // MUFU.RCP ...
    inv = inv * fmaf(-w, inv, 2.f);                   // ... + one Newton step (~1 ulp; two IEEE divisions cost 5x more)
    gx = fmaf(h[0], u, fmaf(h[1], v, h[2])) * inv;
    gy = fmaf(h[3], u, fmaf(h[4], v, h[5])) * inv;
    return w > 0.f && gx > 0.f && gx < (float)qw && gy > 0.f && gy < (float)qh;
}

__device__ __forceinline__ unsigned hit_mask(const CompositeParams& p, float u, float v) {
    unsigned mask = 0u;
    const int qw = p.view.qw, qh = p.view.qh;
    for (int d = 0; d < p.view.D; ++d) {
        float gx, gy;
        if (plane_grid(&p.view.hom[d * 9], u, v, qw, qh, gx, gy)) {
            const int qx = min((int)gx, qw - 1), qy = min((int)gy, qh - 1);
            const int kind = __ldg(&p.quads[(d * qh + qy) * qw + qx].kind);
            if (kind != 0) mask |= (1u << d);
        }
    }
    return mask;
}

// Tap geometry of one (pixel, plane) sample: texel offset of the top-left tap, the four bilinear
// weights (zero for taps outside the atlas: grid_sample padding_mode="zeros", MPV.py:425-427) and
// clamped neighbour offsets.
struct Taps {
    int o00, o10, o01, o11;   // texel offsets (units of float4)
    float w00, w10, w01, w11;
    int kind;
};

__device__ __forceinline__ Taps taps_from_grid(const CompositeParams& p, int d, float gx, float gy) {
    Taps t;
    const int qw = p.view.qw, qh = p.view.qh;
    const int qx = min((int)gx, qw - 1), qy = min((int)gy, qh - 1);
    const float4* qp = reinterpret_cast<const float4*>(&p.quads[(d * qh + qy) * qw + qx]);
    const float4 qa = __ldg(qp);
    const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
    const float a = gx - (float)qx, b = gy - (float)qy;
    const float lx = fmaf(a, qa.z, qa.x), ly = fmaf(b, qa.w, qa.y);
    const float flx = floorf(lx), fly = floorf(ly);
    const float fx = lx - flx, fy = ly - fly;
    const int ix = qb.x + (int)flx, iy = qb.y + (int)fly;            // >= 0: tiles lie inside the atlas (host-checked)
    t.kind = qb.z;
    const int aw = (t.kind == 2) ? p.view.dyn_w : p.view.sta_w;
    const int ah = (t.kind == 2) ? p.view.dyn_h : p.view.sta_h;
    // grid_sample zero padding can only trigger on the last row / column of the atlas
    const bool x0ok = ix < aw, x1ok = ix + 1 < aw, y0ok = iy < ah, y1ok = iy + 1 < ah;
    const int cx0 = min(ix, aw - 1), cx1 = min(ix + 1, aw - 1);
    const int cy0 = min(iy, ah - 1), cy1 = min(iy + 1, ah - 1);
    t.o00 = cy0 * aw + cx0; t.o10 = cy0 * aw + cx1;
    t.o01 = cy1 * aw + cx0; t.o11 = cy1 * aw + cx1;
    const float gx1 = x1ok ? fx : 0.f, gx0 = x0ok ? 1.f - fx : 0.f;
    const float gy1 = y1ok ? fy : 0.f, gy0 = y0ok ? 1.f - fy : 0.f;
    t.w00 = gx0 * gy0; t.w10 = gx1 * gy0; t.w01 = gx0 * gy1; t.w11 = gx1 * gy1;
    return t;
}

__device__ __forceinline__ Taps make_taps(const CompositeParams& p, int d, float u, float v) {
    float gx, gy;
    plane_grid(&p.view.hom[d * 9], u, v, p.view.qw, p.view.qh, gx, gy);
    return taps_from_grid(p, d, gx, gy);
}

__device__ __forceinline__ float4 sample_rgba(const float4* __restrict__ base, const Taps& t) {
    const float4 a = ldg4(base + t.o00), b = ldg4(base + t.o10), c = ldg4(base + t.o01), d = ldg4(base + t.o11);
    float4 r;
    r.x = a.x * t.w00 + b.x * t.w10 + c.x * t.w01 + d.x * t.w11;
    r.y = a.y * t.w00 + b.y * t.w10 + c.y * t.w01 + d.y * t.w11;
    r.z = a.z * t.w00 + b.z * t.w10 + c.z * t.w01 + d.z * t.w11;
    r.w = a.w * t.w00 + b.w * t.w10 + c.w * t.w01 + d.w * t.w11;
    // rgb_activate / alpha_activate = sigmoid (MPV.py:435, MPI.py:22)
    r.x = sigmoidf_fast(r.x); r.y = sigmoidf_fast(r.y); r.z = sigmoidf_fast(r.z); r.w = sigmoidf_fast(r.w);
    return r;
}

__device__ __forceinline__ int block_max(int v, int* sm) {
    v = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) sm[(threadIdx.y * BX + threadIdx.x) >> 5] = v;
    __syncthreads();
    int m = 0;
#pragma unroll
    for (int i = 0; i < (BX * BY) / 32; ++i) m = max(m, sm[i]);
    __syncthreads();
    return m;
}

static inline int validate_view(const vl3d_view* v, const vl3d_quad* quads, const float* dyn, const float* sta) {
    VL3D_REQUIRE(v != nullptr && quads != nullptr, VL3D_ENULL, "view / quads is NULL");
    VL3D_REQUIRE(v->D >= 1 && v->D <= VL3D_MAX_PLANES, VL3D_ERANGE, "D=%d outside [1,%d]", v->D, VL3D_MAX_PLANES);
    VL3D_REQUIRE(v->H >= 1 && v->W >= 1 && v->qh >= 1 && v->qw >= 1, VL3D_EINVAL, "bad view sizes");
    VL3D_REQUIRE(dyn != nullptr || (v->dyn_h == 0 && v->dyn_w == 0), VL3D_ENULL, "atlas_dyn is NULL");
    VL3D_REQUIRE(sta != nullptr || (v->sta_h == 0 && v->sta_w == 0), VL3D_ENULL, "atlas_sta is NULL");
    VL3D_REQUIRE(((uintptr_t)dyn & 15) == 0 && ((uintptr_t)sta & 15) == 0 && ((uintptr_t)quads & 15) == 0,
                 VL3D_EALIGN, "atlas / quad pointers must be 16-byte aligned");
    return 0;
}

}  // namespace vl3d
