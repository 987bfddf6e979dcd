// MPV tile alpha-composite forward / backward for sm_100a.
//
// Replaces (reference file:line): MPV.py:353-405 (rasterise + uv lookup -> analytic per-plane
// homography + quad table), MPV.py:413-449 (grid_sample + sigmoid + masked_scatter),
// utils_mpi.py:92-107 (overcompose), MPV.py:454 (alpha), MPV.py:517-531 (slot-wise smoothness)
// and autograd's backward through all of them.
//
// Kernels: composite_fwd_kernel (below; slot-ordered forward with regulariser sums / debug outputs, used by the
// autograd path), composite_render_kernel / composite_bwd_kernel (composite_lean.cuh) and the TMA-staged render
// (composite_tma.cuh).  The fused backward + Adam pass lives in fused_bwd_adam.cu.
// The dense (T,H,W,K,4) `mpi` tensor of the reference is never materialised (optional debug output).
#include "composite_common.cuh"

namespace vl3d {

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

}  // namespace vl3d

#include "tma_common.cuh"
#include "composite_lean.cuh"
#include "composite_tma.cuh"

namespace vl3d {

// Frames per thread (geometry is shared by the frames of a chunk), chosen on B200 at 720p (profiles/README.md):
// render 3 (TMA: 3 stages), backward 2.  A call is split into a TF-multiple and a tail of single frames.
constexpr int FWD_TF = 3, BWD_TF = 2, RENDER_TMA_STAGES = 3;

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

template <bool SMOOTH>
static void launch_bwd(const CompositeParams& p0, int T, dim3 grid2, dim3 block, cudaStream_t st, const float* atlas_dyn,
                       bool rect_planes) {
    constexpr int TF = BWD_TF;
    TmaRenderParams P;
    P.p = p0;
    const int nz = T / TF;
    if (nz > 0) {
        P.p.tb = 0;
        bool split = false;
        if constexpr (SMOOTH) {
            // dense layout: tiles whose pixels all hit the same planes stage their texels with TMA, the rest (image
            // border) keep the per-thread loads — decided per tile inside one launch
            if (rect_planes && P.p.ts == nullptr && make_atlas_tmap(&P.tmap, P.p.view, atlas_dyn, T)) {
                const size_t smem = (size_t)BWD_TMA_STAGES * TF * TMA_BOX_BYTES;
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
    if (T % TF) {
        P.p.tb = nz * TF;
        composite_bwd_kernel<1, SMOOTH, 0><<<dim3(grid2.x, grid2.y, T % TF), block, 0, st>>>(P);
    }
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
    const int tf = (!smooth && !mpi) ? FWD_TF : 4;
    dim3 grid((view->W + sx - 1) / sx, (view->H + sy - 1) / sy, (T + tf - 1) / tf), block(BX, BY);
    cudaStream_t st = (cudaStream_t)stream;
    if (smooth && mpi) composite_fwd_kernel<4, true, true><<<grid, block, 0, st>>>(p);
    else if (smooth) composite_fwd_kernel<4, true, false><<<grid, block, 0, st>>>(p);
    else if (mpi) composite_fwd_kernel<4, false, true><<<grid, block, 0, st>>>(p);
    else {
        // dense layout (VL3D_VIEW_RECT_PLANES): TMA-staged render (composite_tma.cuh); otherwise per-thread loads
        int main_frames = 0;
        if ((view->flags & VL3D_VIEW_RECT_PLANES) && ts == nullptr && atlas_dyn != nullptr &&
            launch_render_tma<FWD_TF, RENDER_TMA_STAGES>(p, atlas_dyn, T, st))
            main_frames = T / FWD_TF * FWD_TF;
        if (main_frames == 0) {
            launch_render<FWD_TF, 4>(p, T, grid, block, st);
        } else if (main_frames < T) {                               // tail frames: per-thread loads, one frame per CTA
            CompositeParams q = p;
            q.tb = main_frames;
            composite_render_kernel<1, 4><<<dim3(grid.x, grid.y, T - main_frames), block, 0, st>>>(q);
        }
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
    const int sx = smooth ? BX - 1 : BX, sy = smooth ? BY - 1 : BY;
    dim3 grid((view->W + sx - 1) / sx, (view->H + sy - 1) / sy, 1), block(BX, BY);
    cudaStream_t st = (cudaStream_t)stream;
    const bool rect = (view->flags & VL3D_VIEW_RECT_PLANES) != 0;
    if (smooth) launch_bwd<true>(p, T, grid, block, st, atlas_dyn, rect);
    else launch_bwd<false>(p, T, grid, block, st, atlas_dyn, rect);
    return check_launch("composite_bwd");
}
