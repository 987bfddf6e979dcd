// MPV tile alpha-composite forward / backward for sm_100a.
//
// Replaces (reference file:line): MPV.py:353-405 (rasterise + uv lookup -> analytic per-plane
// homography + quad table), MPV.py:413-449 (grid_sample + sigmoid + masked_scatter),
// utils_mpi.py:92-107 (overcompose), MPV.py:454 (alpha), MPV.py:517-531 (slot-wise smoothness)
// and autograd's backward through all of them.
//
// Mapping: one thread = one screen pixel, 32 lanes = 32 adjacent pixels of a row, so the four
// bilinear taps of a warp are four coalesced 512-byte runs of RGBA texels (one LDG.128 per tap).
// A CTA is a 32x8 pixel tile x a chunk of TF frames; geometry (hit mask, tap address, bilinear
// weights) depends on (pixel, plane) only and is shared by the TF frames in registers.
// The loop runs over *slots* (k-th hit along the ray, utils.py:64-69) so neighbouring pixels are
// slot-aligned for the smoothness terms: right neighbour by warp shuffle, lower neighbour through a
// double-buffered shared-memory tile.  With smoothness on, tiles overlap by one pixel column / row
// (31x7 owned pixels) so every pixel pair lives in exactly one CTA.
// The dense (T,H,W,K,4) `mpi` tensor of the reference is never materialised (optional debug output).
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
    int dbg_nored;   // tuning aids (results are then wrong): bit0 skip REDs, bit1 skip sign maths, bit2 skip sums, bit3 skip exchange
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

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int TF, bool SMOOTH, bool MPI>
__global__ void __launch_bounds__(BX* BY) composite_fwd_kernel(const __grid_constant__ CompositeParams p) {
    constexpr int SX = SMOOTH ? BX - 1 : BX, SY = SMOOTH ? BY - 1 : BY;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int px = blockIdx.x * SX + tx, py = blockIdx.y * SY + ty;
    const int H = p.view.H, W = p.view.W;
    const bool active = px < W && py < H;
    const bool owned = active && tx < SX && ty < SY;
    const int t0 = blockIdx.z * TF;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;

    __shared__ int s_red[(BX * BY) / 32];
    __shared__ float4 s_ex[SMOOTH ? 2 : 1][SMOOTH ? TF : 1][SMOOTH ? BY : 1][SMOOTH ? BX : 1];

    unsigned rem = active ? hit_mask(p, u, v) : 0u;
    const int nhit = __popc(rem);
    if (p.hits_out != nullptr && owned && blockIdx.z == 0) p.hits_out[py * W + px] = nhit;
    const int kmax = (SMOOTH || MPI) ? block_max(nhit, s_red) : nhit;

    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* fbase[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = min(t0 + f, p.T - 1);
        const int ft = p.ts ? __ldg(&p.ts[t]) : t;
        fbase[f] = p.atlas_dyn + (size_t)ft * dyn_frame;
    }

    float Tr[TF], cr[TF], cg[TF], cb[TF], ca[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) { Tr[f] = 1.f; cr[f] = cg[f] = cb[f] = ca[f] = 0.f; }
    float sxr = 0.f, syr = 0.f, sxa = 0.f, sya = 0.f;
    const bool pair_x = owned && (px + 1 < W), pair_y = owned && (py + 1 < H);

    for (int k = 0; k < kmax; ++k) {
        const bool has = rem != 0u;
        float4 val[TF];
        if (has) {
            const int d = __ffs(rem) - 1;
            rem &= rem - 1u;
            const Taps tp = make_taps(p, d, u, v);
            if (tp.kind == 2) {
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = sample_rgba(fbase[f], tp);
            } else {
                const float4 s = sample_rgba(p.atlas_sta, tp);   // static tile: same for all frames (MPV.py:445)
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = s;
            }
#pragma unroll
            for (int f = 0; f < TF; ++f) {
                const float bw = val[f].w * Tr[f];                // utils_mpi.py:100-104
                cr[f] = fmaf(bw, val[f].x, cr[f]);
                cg[f] = fmaf(bw, val[f].y, cg[f]);
                cb[f] = fmaf(bw, val[f].z, cb[f]);
                ca[f] += bw;
                Tr[f] *= (1.f - val[f].w);
            }
        } else {
#pragma unroll
            for (int f = 0; f < TF; ++f) val[f] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero canvas (MPV.py:441)
        }
        if (MPI) {
            if (owned) {
#pragma unroll
                for (int f = 0; f < TF; ++f)
                    if (t0 + f < p.T)
                        p.mpi_out[(((size_t)(t0 + f) * H + py) * W + px) * p.view.D + k] = val[f];
            }
        }
        if (SMOOTH) {
            const int buf = k & 1;
#pragma unroll
            for (int f = 0; f < TF; ++f) s_ex[buf][f][ty][tx] = val[f];
            __syncthreads();
#pragma unroll
            for (int f = 0; f < TF; ++f) {
                const bool fok = t0 + f < p.T;
                float4 r;
                r.x = __shfl_down_sync(0xffffffffu, val[f].x, 1);
                r.y = __shfl_down_sync(0xffffffffu, val[f].y, 1);
                r.z = __shfl_down_sync(0xffffffffu, val[f].z, 1);
                r.w = __shfl_down_sync(0xffffffffu, val[f].w, 1);
                if (pair_x && fok) {
                    sxr += fabsf(val[f].x - r.x) + fabsf(val[f].y - r.y) + fabsf(val[f].z - r.z);
                    sxa += fabsf(val[f].w - r.w);
                }
                if (pair_y && fok) {
                    const float4 dn = s_ex[buf][f][ty + 1][tx];
                    syr += fabsf(val[f].x - dn.x) + fabsf(val[f].y - dn.y) + fabsf(val[f].z - dn.z);
                    sya += fabsf(val[f].w - dn.w);
                }
            }
            // the other buffer is rewritten next iteration; its readers finished before this sync
        }
    }

    if (owned) {
        const size_t plane = (size_t)H * W;
        const size_t pix = (size_t)py * W + px;
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            const int t = t0 + f;
            if (t < p.T) {
                float* o = p.rgb_out + (size_t)t * 3 * plane + pix;
                o[0] = cr[f]; o[plane] = cg[f]; o[2 * plane] = cb[f];
                if (t < p.pad) {                                  // loop pad: cat(rgb, rgb[:pt-1]) (MPV.py:490-492)
                    float* o2 = p.rgb_out + (size_t)(p.T + t) * 3 * plane + pix;
                    o2[0] = cr[f]; o2[plane] = cg[f]; o2[2 * plane] = cb[f];
                }
                if (p.alpha_out) p.alpha_out[(size_t)t * plane + pix] = ca[f];
            }
        }
    }
    if (SMOOTH) {
        __shared__ float s_sum[4][(BX * BY) / 32];
        const int warp = (ty * BX + tx) >> 5;
        float a0 = warp_sum(sxr), a1 = warp_sum(syr), a2 = warp_sum(sxa), a3 = warp_sum(sya);
        if (tx == 0) { s_sum[0][warp] = a0; s_sum[1][warp] = a1; s_sum[2][warp] = a2; s_sum[3][warp] = a3; }
        __syncthreads();
        if (ty == 0 && tx < 4) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < (BX * BY) / 32; ++i) acc += (double)s_sum[tx][i];
            atomicAdd(&p.smooth[tx], acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pure render (no regulariser, no debug outputs): planes are visited in lockstep (the plane index is
// CTA-uniform, so its homography lives in uniform registers), the homography is evaluated once per
// (pixel, plane) and no slot bookkeeping is needed because nothing compares neighbouring pixels.
// ------------------------------------------------------------------------------------------------
template <int TF>
__global__ void __launch_bounds__(BX* BY) composite_render_v1_kernel(const __grid_constant__ CompositeParams p) {
    const int px = blockIdx.x * BX + threadIdx.x, py = blockIdx.y * BY + threadIdx.y;
    const int H = p.view.H, W = p.view.W;
    if (px >= W || py >= H) return;
    const int t0 = blockIdx.z * TF;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;
    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* fbase[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = min(t0 + f, p.T - 1);
        const int ft = p.ts ? __ldg(&p.ts[t]) : t;
        fbase[f] = p.atlas_dyn + (size_t)ft * dyn_frame;
    }
    float Tr[TF], cr[TF], cg[TF], cb[TF], ca[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) { Tr[f] = 1.f; cr[f] = cg[f] = cb[f] = ca[f] = 0.f; }
    const int D = p.view.D, qw = p.view.qw, qh = p.view.qh;
    int nhit = 0;
    for (int d = 0; d < D; ++d) {
        float gx, gy;
        if (!plane_grid(&p.view.hom[d * 9], u, v, qw, qh, gx, gy)) continue;
        const Taps tp = taps_from_grid(p, d, gx, gy);
        if (tp.kind == 0) continue;
        ++nhit;
        float4 val[TF];
        if (tp.kind == 2) {
#pragma unroll
            for (int f = 0; f < TF; ++f) val[f] = sample_rgba(fbase[f], tp);
        } else {
            const float4 s = sample_rgba(p.atlas_sta, tp);
#pragma unroll
            for (int f = 0; f < TF; ++f) val[f] = s;
        }
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            const float bw = val[f].w * Tr[f];                    // utils_mpi.py:100-104
            cr[f] = fmaf(bw, val[f].x, cr[f]);
            cg[f] = fmaf(bw, val[f].y, cg[f]);
            cb[f] = fmaf(bw, val[f].z, cb[f]);
            ca[f] += bw;
            Tr[f] *= (1.f - val[f].w);
        }
    }
    if (p.hits_out != nullptr && blockIdx.z == 0) p.hits_out[py * W + px] = nhit;
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = t0 + f;
        if (t < p.T) {
            float* o = p.rgb_out + (size_t)t * 3 * plane + pix;
            o[0] = cr[f]; o[plane] = cg[f]; o[2 * plane] = cb[f];
            if (t < p.pad) {                                      // loop pad: cat(rgb, rgb[:pt-1]) (MPV.py:490-492)
                float* o2 = p.rgb_out + (size_t)(p.T + t) * 3 * plane + pix;
                o2[0] = cr[f]; o2[plane] = cg[f]; o2[2 * plane] = cb[f];
            }
            if (p.alpha_out) p.alpha_out[(size_t)t * plane + pix] = ca[f];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pure render, software-pipelined: while plane d is being blended, the 4*TF texel taps of the next hit
// plane are already in flight as cp.async (LDGSTS.128) into per-thread shared-memory slots, so a warp
// never sits on the global-load latency between geometry and blend.  Slots are private to a thread:
// no barrier, only cp.async.wait_group.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float4* smem_dst, const float4* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct TapW { float w00, w10, w01, w11; int kind; };

template <int TF>
__global__ void __launch_bounds__(BX* BY) composite_render_pipe_kernel(const __grid_constant__ CompositeParams p) {
    extern __shared__ __align__(16) float4 s_tap[];              // [2][TF*4][BX*BY]
    constexpr int NT = BX * BY;
    const int tid = threadIdx.y * BX + threadIdx.x;
    const int px = blockIdx.x * BX + threadIdx.x, py = blockIdx.y * BY + threadIdx.y;
    const int H = p.view.H, W = p.view.W;
    if (px >= W || py >= H) return;                               // no barriers below
    const int t0 = blockIdx.z * TF;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;
    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* fbase[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = min(t0 + f, p.T - 1);
        const int ft = p.ts ? __ldg(&p.ts[t]) : t;
        fbase[f] = p.atlas_dyn + (size_t)ft * dyn_frame;
    }
    float Tr[TF], cr[TF], cg[TF], cb[TF], ca[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) { Tr[f] = 1.f; cr[f] = cg[f] = cb[f] = ca[f] = 0.f; }
    const int D = p.view.D, qw = p.view.qw, qh = p.view.qh;

    // find the next hit plane at or after d, issue its tap loads into buffer `buf`; returns the plane (D if none)
    auto issue = [&](int d, int buf, TapW& tw) -> int {
        for (; d < D; ++d) {
            float gx, gy;
            if (!plane_grid(&p.view.hom[d * 9], u, v, qw, qh, gx, gy)) continue;
            const Taps tp = taps_from_grid(p, d, gx, gy);
            if (tp.kind == 0) continue;
            tw.w00 = tp.w00; tw.w10 = tp.w10; tw.w01 = tp.w01; tw.w11 = tp.w11; tw.kind = tp.kind;
            float4* dst = s_tap + (size_t)buf * (TF * 4) * NT + tid;
            if (tp.kind == 2) {
#pragma unroll
                for (int f = 0; f < TF; ++f) {
                    cp_async16(dst + (f * 4 + 0) * NT, fbase[f] + tp.o00);
                    cp_async16(dst + (f * 4 + 1) * NT, fbase[f] + tp.o10);
                    cp_async16(dst + (f * 4 + 2) * NT, fbase[f] + tp.o01);
                    cp_async16(dst + (f * 4 + 3) * NT, fbase[f] + tp.o11);
                }
            } else {
                cp_async16(dst + 0 * NT, p.atlas_sta + tp.o00);
                cp_async16(dst + 1 * NT, p.atlas_sta + tp.o10);
                cp_async16(dst + 2 * NT, p.atlas_sta + tp.o01);
                cp_async16(dst + 3 * NT, p.atlas_sta + tp.o11);
            }
            break;
        }
        cp_commit();
        return d;
    };

    TapW cur, nxt;
    int nhit = 0, buf = 0;
    int d = issue(0, 0, cur);
    while (d < D) {
        const int dn = issue(d + 1, buf ^ 1, nxt);               // next plane's taps fly while this one is blended
        cp_wait<1>();
        ++nhit;
        const float4* src = s_tap + (size_t)buf * (TF * 4) * NT + tid;
        float4 val[TF];
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            const int ff = (cur.kind == 2) ? f : 0;
            const float4 a = src[(ff * 4 + 0) * NT], b = src[(ff * 4 + 1) * NT], c = src[(ff * 4 + 2) * NT],
                         e = src[(ff * 4 + 3) * NT];
            float4 r;
            r.x = a.x * cur.w00 + b.x * cur.w10 + c.x * cur.w01 + e.x * cur.w11;
            r.y = a.y * cur.w00 + b.y * cur.w10 + c.y * cur.w01 + e.y * cur.w11;
            r.z = a.z * cur.w00 + b.z * cur.w10 + c.z * cur.w01 + e.z * cur.w11;
            r.w = a.w * cur.w00 + b.w * cur.w10 + c.w * cur.w01 + e.w * cur.w11;
            r.x = sigmoidf_fast(r.x); r.y = sigmoidf_fast(r.y); r.z = sigmoidf_fast(r.z); r.w = sigmoidf_fast(r.w);
            val[f] = r;
        }
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            const float bw = val[f].w * Tr[f];                    // utils_mpi.py:100-104
            cr[f] = fmaf(bw, val[f].x, cr[f]);
            cg[f] = fmaf(bw, val[f].y, cg[f]);
            cb[f] = fmaf(bw, val[f].z, cb[f]);
            ca[f] += bw;
            Tr[f] *= (1.f - val[f].w);
        }
        cur = nxt; d = dn; buf ^= 1;
    }
    cp_wait<0>();
    if (p.hits_out != nullptr && blockIdx.z == 0) p.hits_out[py * W + px] = nhit;
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = t0 + f;
        if (t < p.T) {
            float* o = p.rgb_out + (size_t)t * 3 * plane + pix;
            o[0] = cr[f]; o[plane] = cg[f]; o[2 * plane] = cb[f];
            if (t < p.pad) {
                float* o2 = p.rgb_out + (size_t)(p.T + t) * 3 * plane + pix;
                o2[0] = cr[f]; o2[plane] = cg[f]; o2[2 * plane] = cb[f];
            }
            if (p.alpha_out) p.alpha_out[(size_t)t * plane + pix] = ca[f];
        }
    }
}

template <int TF>
static int launch_render_pipe(const CompositeParams& p, dim3 grid, dim3 block, cudaStream_t st) {
    const size_t smem = (size_t)2 * TF * 4 * BX * BY * sizeof(float4);
    cudaError_t ce = cudaFuncSetAttribute(composite_render_pipe_kernel<TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) return set_err((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    composite_render_pipe_kernel<TF><<<grid, block, smem, st>>>(p);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// backward: recompute the forward front-to-back; with total = g . rgb_final saved from the forward,
// the suffix sum S_k = sum_{j>k} bw_j (g.c_j) is total - prefix_k, so one pass suffices:
//   dL/dlogit_c = g * bw_k * c(1-c)
//   dL/dlogit_a = a(1-a) T_k (g.c_k) - a S_k          (Appendix B.2 of SURVEY.md, chain through sigmoid)
// plus the sign-gradients of the smoothness terms (branch-free: pair masks are folded into the four
// direction weights).  The kernel can also emit the smoothness sums themselves, so the fused step's
// forward pass is a pure render.  Texel gradients go out as one RED.128 per tap.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wsign(float w, float d) { return copysignf(d != 0.f ? w : 0.f, d); }   // w * sign(d)

__device__ __forceinline__ void red_tap(float4* base, int off, const float4& g, float w) {
    red_add_v4(base + off, make_float4(g.x * w, g.y * w, g.z * w, g.w * w));
}

template <int TF, bool SMOOTH>
__global__ void __launch_bounds__(BX* BY) composite_bwd_v1_kernel(const __grid_constant__ CompositeParams p) {
    constexpr int SX = SMOOTH ? BX - 1 : BX, SY = SMOOTH ? BY - 1 : BY;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int px = blockIdx.x * SX + tx, py = blockIdx.y * SY + ty;
    const int H = p.view.H, W = p.view.W;
    const bool active = px < W && py < H;
    const bool owned = active && tx < SX && ty < SY;
    const int t0 = blockIdx.z * TF;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;

    __shared__ int s_red[(BX * BY) / 32];
    __shared__ float4 s_ex[SMOOTH ? 2 : 1][SMOOTH ? TF : 1][SMOOTH ? BY : 1][SMOOTH ? BX : 1];

    unsigned rem = active ? hit_mask(p, u, v) : 0u;
    const int nhit = __popc(rem);
    const int kmax = SMOOTH ? block_max(nhit, s_red) : __reduce_max_sync(0xffffffffu, nhit);

    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    size_t foff[TF];
    float g0[TF], g1[TF], g2[TF], tot[TF], Tr[TF], pre[TF];
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = min(t0 + f, p.T - 1);
        const int ft = p.ts ? __ldg(&p.ts[t]) : t;
        foff[f] = (size_t)ft * dyn_frame;
        g0[f] = g1[f] = g2[f] = 0.f; tot[f] = 0.f; Tr[f] = 1.f; pre[f] = 0.f;
        if (owned && t0 + f < p.T) {
            const float* gp = p.grad_rgb + (size_t)t * 3 * plane + pix;
            g0[f] = gp[0]; g1[f] = gp[plane]; g2[f] = gp[2 * plane];
            if (t < p.pad) {                                      // adjoint of cat(rgb, rgb[:pad])
                const float* gq = p.grad_rgb + (size_t)(p.T + t) * 3 * plane + pix;
                g0[f] += gq[0]; g1[f] += gq[plane]; g2[f] += gq[2 * plane];
            }
            const float* rp = p.rgb + (size_t)t * 3 * plane + pix;
            tot[f] = g0[f] * rp[0] + g1[f] * rp[plane] + g2[f] * rp[2 * plane];
        }
    }
    // pair ownership folded into per-direction weights (0 where the pair does not exist / is not ours):
    //   right / down: pairs owned by this pixel; left / up: pairs owned by the neighbour inside this CTA
    float wr_c = 0.f, wl_c = 0.f, wd_c = 0.f, wu_c = 0.f, wr_a = 0.f, wl_a = 0.f, wd_a = 0.f, wu_a = 0.f;
    float mx = 0.f, my = 0.f;
    if (SMOOTH) {
        const float wxr = __ldg(p.w_smooth), wyr = __ldg(p.w_smooth + 1), wxa = __ldg(p.w_smooth + 2), wya = __ldg(p.w_smooth + 3);
        const bool pair_x = owned && (px + 1 < W), pair_y = owned && (py + 1 < H);
        const bool pair_l = active && tx >= 1 && ty < SY;        // left neighbour (tx-1,ty) is owned
        const bool pair_u = active && ty >= 1 && tx < SX;        // upper neighbour (tx,ty-1) is owned
        if (pair_x) { wr_c = wxr; wr_a = wxa; mx = 1.f; }
        if (pair_l) { wl_c = wxr; wl_a = wxa; }
        if (pair_y) { wd_c = wyr; wd_a = wya; my = 1.f; }
        if (pair_u) { wu_c = wyr; wu_a = wya; }
    }
    float sxr = 0.f, syr = 0.f, sxa = 0.f, sya = 0.f;

    for (int k = 0; k < kmax; ++k) {
        const bool has = rem != 0u;
        Taps tp;
        tp.kind = 0; tp.o00 = tp.o10 = tp.o01 = tp.o11 = -1;
        tp.w00 = tp.w10 = tp.w01 = tp.w11 = 0.f;
        float4 val[TF];
#pragma unroll
        for (int f = 0; f < TF; ++f) val[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) {
            const int d = __ffs(rem) - 1;
            rem &= rem - 1u;
            tp = make_taps(p, d, u, v);
            if (tp.kind == 2) {
#pragma unroll
                for (int f = 0; f < TF; ++f)
                    if (t0 + f < p.T) val[f] = sample_rgba(p.atlas_dyn + foff[f], tp);   // frames past T stay 0
            } else {
                const float4 s = sample_rgba(p.atlas_sta, tp);
#pragma unroll
                for (int f = 0; f < TF; ++f)
                    if (t0 + f < p.T) val[f] = s;
            }
        }
        float4 gs[TF];   // dL/d(activated value) from the smoothness terms
#pragma unroll
        for (int f = 0; f < TF; ++f) gs[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SMOOTH && !(p.dbg_nored & 8)) {
            const int buf = k & 1;
#pragma unroll
            for (int f = 0; f < TF; ++f) s_ex[buf][f][ty][tx] = val[f];
            __syncthreads();
            const int tyd = min(ty + 1, BY - 1), tyu = max(ty - 1, 0);
#pragma unroll
            for (int f = 0; f < TF; ++f) {
                const float4 c = val[f];
                float4 r, l;
                r.x = __shfl_down_sync(0xffffffffu, c.x, 1); l.x = __shfl_up_sync(0xffffffffu, c.x, 1);
                r.y = __shfl_down_sync(0xffffffffu, c.y, 1); l.y = __shfl_up_sync(0xffffffffu, c.y, 1);
                r.z = __shfl_down_sync(0xffffffffu, c.z, 1); l.z = __shfl_up_sync(0xffffffffu, c.z, 1);
                r.w = __shfl_down_sync(0xffffffffu, c.w, 1); l.w = __shfl_up_sync(0xffffffffu, c.w, 1);
                const float4 dn = s_ex[buf][f][tyd][tx], up = s_ex[buf][f][tyu][tx];
                const float dxr = c.x - r.x, dyr = c.y - r.y, dzr = c.z - r.z, dwr = c.w - r.w;
                const float dxd = c.x - dn.x, dyd = c.y - dn.y, dzd = c.z - dn.z, dwd = c.w - dn.w;
                // d|a-b|/da = sign(a-b);  the pair (left, this) contributes -sign(left - this) = sign(this - left)
                if (!(p.dbg_nored & 2)) {
                gs[f].x = wsign(wr_c, dxr) + wsign(wl_c, c.x - l.x) + wsign(wd_c, dxd) + wsign(wu_c, c.x - up.x);
                gs[f].y = wsign(wr_c, dyr) + wsign(wl_c, c.y - l.y) + wsign(wd_c, dyd) + wsign(wu_c, c.y - up.y);
                gs[f].z = wsign(wr_c, dzr) + wsign(wl_c, c.z - l.z) + wsign(wd_c, dzd) + wsign(wu_c, c.z - up.z);
                gs[f].w = wsign(wr_a, dwr) + wsign(wl_a, c.w - l.w) + wsign(wd_a, dwd) + wsign(wu_a, c.w - up.w);
                } else { gs[f].x = dxr + l.x + up.x; gs[f].y = dyd; gs[f].z = dzr; gs[f].w = dwd + dwr; }
                if (p.smooth != nullptr && !(p.dbg_nored & 4)) {                         // the regulariser values themselves (MPV.py:517-531)
                    sxr = fmaf(mx, fabsf(dxr) + fabsf(dyr) + fabsf(dzr), sxr);
                    sxa = fmaf(mx, fabsf(dwr), sxa);
                    syr = fmaf(my, fabsf(dxd) + fabsf(dyd) + fabsf(dzd), syr);
                    sya = fmaf(my, fabsf(dwd), sya);
                }
            }
        }
        // ---- gradients of this slot.  The LSU retires roughly one RED lane per 1.3 cycles, so 4 RED.128 per
        // sample would bound the kernel; neighbouring lanes' bilinear footprints overlap (lane i's right
        // taps are usually lane i+1's left taps), so the right-hand contributions are handed to the next
        // lane by shuffle and folded into its left-hand RED: ~2 REDs per sample instead of 4.
        const int kindv = has ? tp.kind : 0;
        const int up_kind = __shfl_up_sync(0xffffffffu, kindv, 1), dn_kind = __shfl_down_sync(0xffffffffu, kindv, 1);
        const int up_o10 = __shfl_up_sync(0xffffffffu, tp.o10, 1), up_o11 = __shfl_up_sync(0xffffffffu, tp.o11, 1);
        const int dn_o00 = __shfl_down_sync(0xffffffffu, tp.o00, 1), dn_o01 = __shfl_down_sync(0xffffffffu, tp.o01, 1);
        const bool recv0 = tx > 0 && kindv != 0 && up_kind == kindv && up_o10 == tp.o00;       // absorb lane-1's right/top tap
        const bool recv1 = tx > 0 && kindv != 0 && up_kind == kindv && up_o11 == tp.o01;
        const bool sent0 = tx < BX - 1 && kindv != 0 && dn_kind == kindv && dn_o00 == tp.o10;  // lane+1 absorbs mine
        const bool sent1 = tx < BX - 1 && kindv != 0 && dn_kind == kindv && dn_o01 == tp.o11;
        float4 gsta = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);   // gradient w.r.t. the pre-sigmoid bilinear sample
            if (has && t0 + f < p.T) {
                const float4 c = val[f];
                const float a = c.w, om = 1.f - a;
                const float bw = a * Tr[f];
                const float gc = g0[f] * c.x + g1[f] * c.y + g2[f] * c.z;
                pre[f] = fmaf(bw, gc, pre[f]);
                const float S = tot[f] - pre[f];
                gl.x = fmaf(g0[f], bw, gs[f].x) * (c.x - c.x * c.x);
                gl.y = fmaf(g1[f], bw, gs[f].y) * (c.y - c.y * c.y);
                gl.z = fmaf(g2[f], bw, gs[f].z) * (c.z - c.z * c.z);
                gl.w = a * (om * fmaf(Tr[f], gc, gs[f].w) - S);
                Tr[f] *= om;
            }
            if (kindv == 1) { gsta.x += gl.x; gsta.y += gl.y; gsta.z += gl.z; gsta.w += gl.w; }   // sum over frames (MPV.py:445)
            const bool dynf = kindv == 2 && t0 + f < p.T;
            float4 l0 = make_float4(gl.x * tp.w00, gl.y * tp.w00, gl.z * tp.w00, gl.w * tp.w00);
            float4 l1 = make_float4(gl.x * tp.w01, gl.y * tp.w01, gl.z * tp.w01, gl.w * tp.w01);
            const float4 r0 = make_float4(gl.x * tp.w10, gl.y * tp.w10, gl.z * tp.w10, gl.w * tp.w10);
            const float4 r1 = make_float4(gl.x * tp.w11, gl.y * tp.w11, gl.z * tp.w11, gl.w * tp.w11);
            float4 i0, i1;
            i0.x = __shfl_up_sync(0xffffffffu, r0.x, 1); i0.y = __shfl_up_sync(0xffffffffu, r0.y, 1);
            i0.z = __shfl_up_sync(0xffffffffu, r0.z, 1); i0.w = __shfl_up_sync(0xffffffffu, r0.w, 1);
            i1.x = __shfl_up_sync(0xffffffffu, r1.x, 1); i1.y = __shfl_up_sync(0xffffffffu, r1.y, 1);
            i1.z = __shfl_up_sync(0xffffffffu, r1.z, 1); i1.w = __shfl_up_sync(0xffffffffu, r1.w, 1);
            if (recv0) { l0.x += i0.x; l0.y += i0.y; l0.z += i0.z; l0.w += i0.w; }
            if (recv1) { l1.x += i1.x; l1.y += i1.y; l1.z += i1.z; l1.w += i1.w; }
            if (dynf && !(p.dbg_nored & 1)) {
                float4* gb = p.grad_dyn + foff[f];
                red_add_v4(gb + tp.o00, l0);
                red_add_v4(gb + tp.o01, l1);
                if (!sent0) red_add_v4(gb + tp.o10, r0);
                if (!sent1) red_add_v4(gb + tp.o11, r1);
            }
        }
        {   // static tiles: one (combined) RED set per slot
            float4 l0 = make_float4(gsta.x * tp.w00, gsta.y * tp.w00, gsta.z * tp.w00, gsta.w * tp.w00);
            float4 l1 = make_float4(gsta.x * tp.w01, gsta.y * tp.w01, gsta.z * tp.w01, gsta.w * tp.w01);
            const float4 r0 = make_float4(gsta.x * tp.w10, gsta.y * tp.w10, gsta.z * tp.w10, gsta.w * tp.w10);
            const float4 r1 = make_float4(gsta.x * tp.w11, gsta.y * tp.w11, gsta.z * tp.w11, gsta.w * tp.w11);
            const bool any_static = __any_sync(0xffffffffu, kindv == 1);
            if (any_static) {
                float4 i0, i1;
                i0.x = __shfl_up_sync(0xffffffffu, r0.x, 1); i0.y = __shfl_up_sync(0xffffffffu, r0.y, 1);
                i0.z = __shfl_up_sync(0xffffffffu, r0.z, 1); i0.w = __shfl_up_sync(0xffffffffu, r0.w, 1);
                i1.x = __shfl_up_sync(0xffffffffu, r1.x, 1); i1.y = __shfl_up_sync(0xffffffffu, r1.y, 1);
                i1.z = __shfl_up_sync(0xffffffffu, r1.z, 1); i1.w = __shfl_up_sync(0xffffffffu, r1.w, 1);
                if (kindv == 1) {
                    if (recv0) { l0.x += i0.x; l0.y += i0.y; l0.z += i0.z; l0.w += i0.w; }
                    if (recv1) { l1.x += i1.x; l1.y += i1.y; l1.z += i1.z; l1.w += i1.w; }
                    red_add_v4(p.grad_sta + tp.o00, l0);
                    red_add_v4(p.grad_sta + tp.o01, l1);
                    if (!sent0) red_add_v4(p.grad_sta + tp.o10, r0);
                    if (!sent1) red_add_v4(p.grad_sta + tp.o11, r1);
                }
            }
        }
    }
    if (SMOOTH && p.smooth != nullptr) {
        __shared__ float s_sum[4][(BX * BY) / 32];
        const int warp = (ty * BX + tx) >> 5;
        const float a0 = warp_sum(sxr), a1 = warp_sum(syr), a2 = warp_sum(sxa), a3 = warp_sum(sya);
        if (tx == 0) { s_sum[0][warp] = a0; s_sum[1][warp] = a1; s_sum[2][warp] = a2; s_sum[3][warp] = a3; }
        __syncthreads();
        if (ty == 0 && tx < 4) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < (BX * BY) / 32; ++i) acc += (double)s_sum[tx][i];
            atomicAdd(&p.smooth[tx], acc);
        }
    }
}

}  // namespace vl3d

#include "tma_common.cuh"
#include "composite_lean.cuh"
#include "composite_tma.cuh"

namespace vl3d {

// frames per thread (geometry is shared by the frames of a chunk).  VL3D_FWD_TF / VL3D_BWD_TF override
// (tuning aids); VL3D_COMPOSITE_V1=1 selects the first-generation kernels (kept for A/B measurements).
static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// (read on every call: getenv costs nanoseconds and in-process sweeps can flip the knobs)
static bool use_v1() { return env_int("VL3D_COMPOSITE_V1", 0) != 0; }

static int fwd_tf(int T) {
    const int env = env_int("VL3D_FWD_TF", 0);
    if (env >= 1 && env <= 4) return env;
    return T >= 3 ? 3 : T;
}

static bool render_pipe() { return env_int("VL3D_PIPE", 0) != 0; }

static int bwd_tf(int T) {
    const int env = env_int("VL3D_BWD_TF", 0);
    if (env >= 1 && env <= 4) return env;
    return T >= 2 ? 2 : 1;
}

// lean kernels: frames [0, T) = nz chunks of TF frames + a tail of T % TF single frames
template <int TF, int MINB>
static void launch_render(CompositeParams p, int T, dim3 grid2, dim3 block, cudaStream_t st) {
    const int nz = T / TF;
    if (nz > 0) {
        p.tb = 0;
        composite_render_kernel<TF, MINB><<<dim3(grid2.x, grid2.y, nz), block, 0, st>>>(p);
    }
    if (TF > 1 && T % TF) {
        p.tb = nz * TF;
        composite_render_kernel<1, 4><<<dim3(grid2.x, grid2.y, T % TF), block, 0, st>>>(p);
    }
}

template <int TF, bool SMOOTH>
static void launch_bwd(const CompositeParams& p0, int T, dim3 grid2, dim3 block, cudaStream_t st, const float* atlas_dyn,
                       bool rect_planes) {
    TmaRenderParams P;
    P.p = p0;
    const int nz = T / TF;
    if (nz > 0) {
        P.p.tb = 0;
        bool split = false;
        if constexpr (SMOOTH && (TF == 2 || TF == 3)) {   // (TF = 3 only through VL3D_BWD_TF=3: round-2 tuning aid)
            // dense layout: tiles whose pixels all hit the same planes stage their texels with TMA, the rest (image
            // border) keep the per-thread loads — decided per tile inside one launch; VL3D_TMA_BWD=0: tuning aid
            if (rect_planes && P.p.ts == nullptr && env_int("VL3D_TMA_BWD", 1) != 0 && make_atlas_tmap(&P.tmap, P.p.view, atlas_dyn, T)) {
                const size_t smem = (size_t)3 * TF * TMA_BOX_BYTES;
                if (cudaFuncSetAttribute(composite_bwd_kernel<TF, SMOOTH, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) ==
                    cudaSuccess) {
                    composite_bwd_kernel<TF, SMOOTH, 3><<<dim3(grid2.x, grid2.y, nz), block, smem, st>>>(P);
                    split = true;
                } else {
                    (void)cudaGetLastError();
                }
            }
        }
        if (!split) composite_bwd_kernel<TF, SMOOTH, 0><<<dim3(grid2.x, grid2.y, nz), block, 0, st>>>(P);
    }
    if (TF > 1 && T % TF) {
        P.p.tb = nz * TF;
        composite_bwd_kernel<1, SMOOTH, 0><<<dim3(grid2.x, grid2.y, T % TF), block, 0, st>>>(P);
    }
}

static int validate_view(const vl3d_view* v, const vl3d_quad* quads, const float* dyn, const float* sta) {
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

using namespace vl3d;

extern "C" int vl3d_composite_fwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                                  const float* atlas_sta, const int32_t* ts, int32_t T, int32_t pad,
                                  float* rgb_out, float* alpha_out, double* smooth_sums, float* mpi_out,
                                  int32_t* hits_out, void* stream) {
    if (int e = validate_view(view, quads, atlas_dyn, atlas_sta)) return e;
    VL3D_REQUIRE(rgb_out != nullptr, VL3D_ENULL, "rgb_out is NULL");
    VL3D_REQUIRE(T >= 1 && pad >= 0 && pad <= T, VL3D_EINVAL, "bad T=%d pad=%d", T, pad);
    VL3D_REQUIRE(((uintptr_t)mpi_out & 15) == 0, VL3D_EALIGN, "mpi_out must be 16-byte aligned");
    CompositeParams p{};
    p.view = *view; p.quads = quads;
    p.atlas_dyn = reinterpret_cast<const float4*>(atlas_dyn);
    p.atlas_sta = reinterpret_cast<const float4*>(atlas_sta);
    p.ts = ts; p.T = T; p.pad = pad;
    p.rgb_out = rgb_out; p.alpha_out = alpha_out; p.smooth = smooth_sums;
    p.mpi_out = reinterpret_cast<float4*>(mpi_out); p.hits_out = hits_out;
    const bool smooth = smooth_sums != nullptr, mpi = mpi_out != nullptr;
    const int sx = smooth ? BX - 1 : BX, sy = smooth ? BY - 1 : BY;
    int tf = (!smooth && !mpi) ? fwd_tf(T) : 4;
    if (!smooth && !mpi && (render_pipe() || use_v1()) && tf > 4) tf = 4;
    dim3 grid((view->W + sx - 1) / sx, (view->H + sy - 1) / sy, (T + tf - 1) / tf), block(BX, BY);
    cudaStream_t st = (cudaStream_t)stream;
    if (smooth && mpi) composite_fwd_kernel<4, true, true><<<grid, block, 0, st>>>(p);
    else if (smooth) composite_fwd_kernel<4, true, false><<<grid, block, 0, st>>>(p);
    else if (mpi) composite_fwd_kernel<4, false, true><<<grid, block, 0, st>>>(p);
    else if (render_pipe()) {
        int e = tf == 1 ? launch_render_pipe<1>(p, grid, block, st) : tf == 2 ? launch_render_pipe<2>(p, grid, block, st)
              : tf == 3 ? launch_render_pipe<3>(p, grid, block, st) : launch_render_pipe<4>(p, grid, block, st);
        if (e) return e;
    }
    else if (use_v1()) {
        if (tf == 3) composite_render_v1_kernel<3><<<grid, block, 0, st>>>(p);
        else if (tf == 2) composite_render_v1_kernel<2><<<grid, block, 0, st>>>(p);
        else if (tf == 1) composite_render_v1_kernel<1><<<grid, block, 0, st>>>(p);
        else composite_render_v1_kernel<4><<<grid, block, 0, st>>>(p);
    } else {
        // MINB = resident CTAs per SM the register budget is capped for (VL3D_FWD_MINB: tuning aid)
        const int minb = env_int("VL3D_FWD_MINB", 0);
        // dense layout: TMA-staged render (composite_tma.cuh); VL3D_TMA=0 selects the per-thread loads (tuning aid)
        bool done = false;
        if ((view->flags & VL3D_VIEW_RECT_PLANES) && ts == nullptr && atlas_dyn != nullptr && env_int("VL3D_TMA", 1) != 0) {
            const int ttf = env_int("VL3D_TMA_TF", 3), tst = env_int("VL3D_TMA_STAGES", 3);   // tuning aids
            int main_frames = 0;
            if (ttf == 2) { if (tst == 4 ? launch_render_tma<2, 4>(p, atlas_dyn, T, st) : launch_render_tma<2, 3>(p, atlas_dyn, T, st)) main_frames = T / 2 * 2; }
            else if (ttf == 4) { if (tst == 2 ? launch_render_tma<4, 2>(p, atlas_dyn, T, st) : launch_render_tma<4, 3>(p, atlas_dyn, T, st)) main_frames = T / 4 * 4; }
            else if (tst == 2 ? launch_render_tma<3, 2>(p, atlas_dyn, T, st)
                     : tst == 4 ? launch_render_tma<3, 4>(p, atlas_dyn, T, st) : launch_render_tma<3, 3>(p, atlas_dyn, T, st)) main_frames = T / 3 * 3;
            if (main_frames > 0) {
                done = true;
                if (main_frames < T) {                              // tail frames: per-thread loads, one frame per CTA
                    CompositeParams q = p;
                    q.tb = main_frames;
                    composite_render_kernel<1, 4><<<dim3(grid.x, grid.y, T - main_frames), block, 0, st>>>(q);
                }
            }
        }
        if (done) {}
        else if (tf == 4) { if (minb == 4) launch_render<4, 4>(p, T, grid, block, st); else launch_render<4, 3>(p, T, grid, block, st); }
        else if (tf == 3) launch_render<3, 4>(p, T, grid, block, st);
        else if (tf == 2) launch_render<2, 4>(p, T, grid, block, st);
        else launch_render<1, 4>(p, T, grid, block, st);
    }
    return check_launch("composite_fwd");
}

extern "C" int vl3d_composite_bwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                                  const float* atlas_sta, const int32_t* ts, int32_t T, int32_t pad,
                                  const float* grad_rgb, const float* rgb, const float* w_smooth,
                                  double* smooth_sums, float* grad_dyn, float* grad_sta, void* stream) {
    if (int e = validate_view(view, quads, atlas_dyn, atlas_sta)) return e;
    VL3D_REQUIRE(grad_rgb != nullptr && rgb != nullptr, VL3D_ENULL, "grad_rgb / rgb is NULL");
    VL3D_REQUIRE(T >= 1 && pad >= 0 && pad <= T, VL3D_EINVAL, "bad T=%d pad=%d", T, pad);
    VL3D_REQUIRE(grad_dyn != nullptr || atlas_dyn == nullptr, VL3D_ENULL, "grad_dyn is NULL");
    VL3D_REQUIRE(grad_sta != nullptr || atlas_sta == nullptr, VL3D_ENULL, "grad_sta is NULL");
    VL3D_REQUIRE(((uintptr_t)grad_dyn & 15) == 0 && ((uintptr_t)grad_sta & 15) == 0, VL3D_EALIGN,
                 "gradient pointers must be 16-byte aligned");
    CompositeParams p{};
    p.view = *view; p.quads = quads;
    p.atlas_dyn = reinterpret_cast<const float4*>(atlas_dyn);
    p.atlas_sta = reinterpret_cast<const float4*>(atlas_sta);
    p.ts = ts; p.T = T; p.pad = pad;
    p.grad_rgb = grad_rgb; p.rgb = rgb;
    p.grad_dyn = reinterpret_cast<float4*>(grad_dyn); p.grad_sta = reinterpret_cast<float4*>(grad_sta);
    const bool smooth = w_smooth != nullptr;
    VL3D_REQUIRE(smooth || smooth_sums == nullptr, VL3D_EINVAL, "smooth_sums needs w_smooth");
    p.w_smooth = w_smooth;
    p.smooth = smooth_sums;
    int tf = bwd_tf(T);
    if (use_v1() && tf > 2) tf = 2;
    p.dbg_nored = env_int("VL3D_BWD_NORED", 0);
    const int sx = smooth ? BX - 1 : BX, sy = smooth ? BY - 1 : BY;
    dim3 grid((view->W + sx - 1) / sx, (view->H + sy - 1) / sy, (T + tf - 1) / tf), block(BX, BY);
    cudaStream_t st = (cudaStream_t)stream;
    if (use_v1()) {
        if (smooth) {
            if (tf == 1) composite_bwd_v1_kernel<1, true><<<grid, block, 0, st>>>(p);
            else composite_bwd_v1_kernel<2, true><<<grid, block, 0, st>>>(p);
        } else {
            if (tf == 1) composite_bwd_v1_kernel<1, false><<<grid, block, 0, st>>>(p);
            else composite_bwd_v1_kernel<2, false><<<grid, block, 0, st>>>(p);
        }
    } else if (smooth) {
        if (tf == 1) launch_bwd<1, true>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
        else if (tf == 2) launch_bwd<2, true>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
        else if (tf == 3) launch_bwd<3, true>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
        else launch_bwd<4, true>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
    } else {
        if (tf == 1) launch_bwd<1, false>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
        else if (tf == 2) launch_bwd<2, false>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
        else if (tf == 3) launch_bwd<3, false>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
        else launch_bwd<4, false>(p, T, grid, block, st, atlas_dyn, (view->flags & VL3D_VIEW_RECT_PLANES) != 0);
    }
    return check_launch("composite_bwd");
}
