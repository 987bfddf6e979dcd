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
    if (!(w > 0.f)) return false;
    gx = fmaf(h[0], u, fmaf(h[1], v, h[2])) / w;
    gy = fmaf(h[3], u, fmaf(h[4], v, h[5])) / w;
    return gx > 0.f && gx < (float)qw && gy > 0.f && gy < (float)qh;
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

__device__ __forceinline__ Taps make_taps(const CompositeParams& p, int d, float u, float v) {
    Taps t;
    const int qw = p.view.qw, qh = p.view.qh;
    float gx, gy;
    plane_grid(&p.view.hom[d * 9], u, v, qw, qh, gx, gy);
    const int qx = min((int)gx, qw - 1), qy = min((int)gy, qh - 1);
    const float4* qp = reinterpret_cast<const float4*>(&p.quads[(d * qh + qy) * qw + qx]);
    const float4 qa = __ldg(qp);
    const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
    const float a = gx - (float)qx, b = gy - (float)qy;
    const float lx = fmaf(a, qa.z, qa.x), ly = fmaf(b, qa.w, qa.y);
    const float flx = floorf(lx), fly = floorf(ly);
    const float fx = lx - flx, fy = ly - fly;
    const int ix = qb.x + (int)flx, iy = qb.y + (int)fly;
    t.kind = qb.z;
    const int aw = (t.kind == 2) ? p.view.dyn_w : p.view.sta_w;
    const int ah = (t.kind == 2) ? p.view.dyn_h : p.view.sta_h;
    const bool x0ok = ix >= 0 && ix < aw, x1ok = ix + 1 >= 0 && ix + 1 < aw;
    const bool y0ok = iy >= 0 && iy < ah, y1ok = iy + 1 >= 0 && iy + 1 < ah;
    const int cx0 = min(max(ix, 0), aw - 1), cx1 = min(max(ix + 1, 0), aw - 1);
    const int cy0 = min(max(iy, 0), ah - 1), cy1 = min(max(iy + 1, 0), ah - 1);
    t.o00 = cy0 * aw + cx0; t.o10 = cy0 * aw + cx1;
    t.o01 = cy1 * aw + cx0; t.o11 = cy1 * aw + cx1;
    t.w00 = (x0ok && y0ok) ? (1.f - fx) * (1.f - fy) : 0.f;
    t.w10 = (x1ok && y0ok) ? fx * (1.f - fy) : 0.f;
    t.w01 = (x0ok && y1ok) ? (1.f - fx) * fy : 0.f;
    t.w11 = (x1ok && y1ok) ? fx * fy : 0.f;
    return t;
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
// backward: recompute the forward front-to-back; with total = g . rgb_final saved from the forward,
// the suffix sum S_k = sum_{j>k} bw_j (g.c_j) is total - prefix_k, so one pass suffices:
//   dL/dlogit_c = g * bw_k * c(1-c)
//   dL/dlogit_a = a(1-a) T_k (g.c_k) - a S_k          (Appendix B.2 of SURVEY.md, chain through sigmoid)
// plus the sign-gradients of the smoothness terms.  Texel gradients go out as one RED.128 per tap.
// ------------------------------------------------------------------------------------------------
template <int TF, bool SMOOTH>
__global__ void __launch_bounds__(BX* BY) composite_bwd_kernel(const __grid_constant__ CompositeParams p) {
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
    const int kmax = SMOOTH ? block_max(nhit, s_red) : nhit;

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
    const bool pair_x = owned && (px + 1 < W), pair_y = owned && (py + 1 < H);
    // pairs owned by the left / upper neighbour inside this CTA
    const bool pair_l = SMOOTH && active && tx >= 1 && ty < SY;   // left neighbour (tx-1,ty) is owned
    const bool pair_u = SMOOTH && active && ty >= 1 && tx < SX;   // upper neighbour (tx,ty-1) is owned
    float wxr = 0.f, wyr = 0.f, wxa = 0.f, wya = 0.f;
    if (SMOOTH) { wxr = __ldg(p.w_smooth); wyr = __ldg(p.w_smooth + 1); wxa = __ldg(p.w_smooth + 2); wya = __ldg(p.w_smooth + 3); }

    for (int k = 0; k < kmax; ++k) {
        const bool has = rem != 0u;
        Taps tp;
        float4 val[TF];
        if (has) {
            const int d = __ffs(rem) - 1;
            rem &= rem - 1u;
            tp = make_taps(p, d, u, v);
            if (tp.kind == 2) {
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = sample_rgba(p.atlas_dyn + foff[f], tp);
            } else {
                const float4 s = sample_rgba(p.atlas_sta, tp);
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = s;
            }
        } else {
            tp.kind = 0;
#pragma unroll
            for (int f = 0; f < TF; ++f) val[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 gs[TF];   // dL/d(activated value) from the smoothness terms
#pragma unroll
        for (int f = 0; f < TF; ++f) gs[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SMOOTH) {
            const int buf = k & 1;
#pragma unroll
            for (int f = 0; f < TF; ++f) s_ex[buf][f][ty][tx] = val[f];
            __syncthreads();
#pragma unroll
            for (int f = 0; f < TF; ++f) {
                const bool fok = t0 + f < p.T;
                float4 r, l;
                r.x = __shfl_down_sync(0xffffffffu, val[f].x, 1); l.x = __shfl_up_sync(0xffffffffu, val[f].x, 1);
                r.y = __shfl_down_sync(0xffffffffu, val[f].y, 1); l.y = __shfl_up_sync(0xffffffffu, val[f].y, 1);
                r.z = __shfl_down_sync(0xffffffffu, val[f].z, 1); l.z = __shfl_up_sync(0xffffffffu, val[f].z, 1);
                r.w = __shfl_down_sync(0xffffffffu, val[f].w, 1); l.w = __shfl_up_sync(0xffffffffu, val[f].w, 1);
                if (pair_x && fok) {
                    gs[f].x += wxr * signf(val[f].x - r.x); gs[f].y += wxr * signf(val[f].y - r.y);
                    gs[f].z += wxr * signf(val[f].z - r.z); gs[f].w += wxa * signf(val[f].w - r.w);
                }
                if (pair_l && fok) {
                    gs[f].x -= wxr * signf(l.x - val[f].x); gs[f].y -= wxr * signf(l.y - val[f].y);
                    gs[f].z -= wxr * signf(l.z - val[f].z); gs[f].w -= wxa * signf(l.w - val[f].w);
                }
                if (pair_y && fok) {
                    const float4 dn = s_ex[buf][f][ty + 1][tx];
                    gs[f].x += wyr * signf(val[f].x - dn.x); gs[f].y += wyr * signf(val[f].y - dn.y);
                    gs[f].z += wyr * signf(val[f].z - dn.z); gs[f].w += wya * signf(val[f].w - dn.w);
                }
                if (pair_u && fok) {
                    const float4 up = s_ex[buf][f][ty - 1][tx];
                    gs[f].x -= wyr * signf(up.x - val[f].x); gs[f].y -= wyr * signf(up.y - val[f].y);
                    gs[f].z -= wyr * signf(up.z - val[f].z); gs[f].w -= wya * signf(up.w - val[f].w);
                }
            }
        }
        if (has) {
            float4 gsta = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int f = 0; f < TF; ++f) {
                const float4 c = val[f];
                const float a = c.w, om = 1.f - a;
                const float bw = a * Tr[f];
                const float gc = g0[f] * c.x + g1[f] * c.y + g2[f] * c.z;
                pre[f] = fmaf(bw, gc, pre[f]);
                const float S = tot[f] - pre[f];
                float4 gl;   // gradient w.r.t. the pre-sigmoid bilinear sample
                gl.x = (g0[f] * bw + gs[f].x) * c.x * (1.f - c.x);
                gl.y = (g1[f] * bw + gs[f].y) * c.y * (1.f - c.y);
                gl.z = (g2[f] * bw + gs[f].z) * c.z * (1.f - c.z);
                gl.w = a * (om * (Tr[f] * gc + gs[f].w) - S);
                Tr[f] *= om;
                if (t0 + f < p.T) {
                    if (tp.kind == 2) {
                        float4* gb = p.grad_dyn + foff[f];
                        if (tp.w00 != 0.f) red_add_v4(gb + tp.o00, make_float4(gl.x * tp.w00, gl.y * tp.w00, gl.z * tp.w00, gl.w * tp.w00));
                        if (tp.w10 != 0.f) red_add_v4(gb + tp.o10, make_float4(gl.x * tp.w10, gl.y * tp.w10, gl.z * tp.w10, gl.w * tp.w10));
                        if (tp.w01 != 0.f) red_add_v4(gb + tp.o01, make_float4(gl.x * tp.w01, gl.y * tp.w01, gl.z * tp.w01, gl.w * tp.w01));
                        if (tp.w11 != 0.f) red_add_v4(gb + tp.o11, make_float4(gl.x * tp.w11, gl.y * tp.w11, gl.z * tp.w11, gl.w * tp.w11));
                    } else {
                        gsta.x += gl.x; gsta.y += gl.y; gsta.z += gl.z; gsta.w += gl.w;   // sum over frames (MPV.py:445 expand)
                    }
                }
            }
            if (tp.kind == 1) {
                float4* gb = p.grad_sta;
                if (tp.w00 != 0.f) red_add_v4(gb + tp.o00, make_float4(gsta.x * tp.w00, gsta.y * tp.w00, gsta.z * tp.w00, gsta.w * tp.w00));
                if (tp.w10 != 0.f) red_add_v4(gb + tp.o10, make_float4(gsta.x * tp.w10, gsta.y * tp.w10, gsta.z * tp.w10, gsta.w * tp.w10));
                if (tp.w01 != 0.f) red_add_v4(gb + tp.o01, make_float4(gsta.x * tp.w01, gsta.y * tp.w01, gsta.z * tp.w01, gsta.w * tp.w01));
                if (tp.w11 != 0.f) red_add_v4(gb + tp.o11, make_float4(gsta.x * tp.w11, gsta.y * tp.w11, gsta.z * tp.w11, gsta.w * tp.w11));
            }
        }
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
    constexpr int TF = 4;
    const bool smooth = smooth_sums != nullptr, mpi = mpi_out != nullptr;
    const int sx = smooth ? BX - 1 : BX, sy = smooth ? BY - 1 : BY;
    dim3 grid((view->W + sx - 1) / sx, (view->H + sy - 1) / sy, (T + TF - 1) / TF), block(BX, BY);
    cudaStream_t st = (cudaStream_t)stream;
    if (smooth && mpi) composite_fwd_kernel<TF, true, true><<<grid, block, 0, st>>>(p);
    else if (smooth) composite_fwd_kernel<TF, true, false><<<grid, block, 0, st>>>(p);
    else if (mpi) composite_fwd_kernel<TF, false, true><<<grid, block, 0, st>>>(p);
    else composite_fwd_kernel<TF, false, false><<<grid, block, 0, st>>>(p);
    return check_launch("composite_fwd");
}

extern "C" int vl3d_composite_bwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                                  const float* atlas_sta, const int32_t* ts, int32_t T, int32_t pad,
                                  const float* grad_rgb, const float* rgb, const float* w_smooth,
                                  float* grad_dyn, float* grad_sta, void* stream) {
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
    p.w_smooth = w_smooth;
    constexpr int TF = 4;
    const int sx = smooth ? BX - 1 : BX, sy = smooth ? BY - 1 : BY;
    dim3 grid((view->W + sx - 1) / sx, (view->H + sy - 1) / sy, (T + TF - 1) / TF), block(BX, BY);
    cudaStream_t st = (cudaStream_t)stream;
    if (smooth) composite_bwd_kernel<TF, true><<<grid, block, 0, st>>>(p);
    else composite_bwd_kernel<TF, false><<<grid, block, 0, st>>>(p);
    return check_launch("composite_bwd");
}
