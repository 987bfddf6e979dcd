// Optional per-ray terms of the MPV composite and their backward, for sm_100a.
//
// Replaces (reference file:line): MPV.py:454 `alpha = blend_weight.sum(-1)` as a DIFFERENTIABLE output (needed by the
// background blend MPV.py:455-461 and the density regulariser MPV.py:533-536), MPV.py:384-385,463-464
// `disp = (1 / zbuf * blend_weight).sum(-1)` (d_smooth, MPV.py:538-551) and the sparsity regulariser MPV.py:511-515
// `mean(|a|_1 / max(|a|_2, 1e-4))` over the slot-indexed alpha, plus autograd's backward through all of them.
// The same terms serve the stage-1 model (MPI.py:552-566,603-607,647-650: sparsity with 1e-6, normalised disparity —
// an affine map of 1 / zbuf that the host folds into `inv_depth`).
//
// Every shipped stage-2 config leaves these terms off (weight 0 / bg_color ''), so they are NOT on the step's hot path:
// the kernels below favour being obviously right over speed (one thread per (pixel, frame), scalar float atomics).
// All three terms only produce gradients for the ALPHA logit of the tapped texels; those are accumulated into the
// same grad_dyn / grad_sta buffers the main backward (vl3d_composite_bwd) uses.
#include "composite_common.cuh"

namespace vl3d {

struct TermsParams {
    CompositeParams p;                        // view, quads, atlases, ts, T
    float inv_depth[VL3D_MAX_PLANES * 3];     // per plane: 1 / view depth = a * u + b * v + c  (u, v as in the view)
    float sparsity_eps;
    // forward
    float* alpha_out;                         // (T,H,W) or NULL
    float* disp_out;                          // (T,H,W) or NULL
    double* sparsity_sum;                     // accumulated, or NULL
    // backward
    const float* g_alpha;                     // (T,H,W) or NULL
    const float* g_disp;                      // (T,H,W) or NULL
    const float* w_sparsity;                  // device float: dL/d(sparsity_sum), or NULL
};

struct RaySums {
    float A, Dp, L1, L2;
};

__device__ __forceinline__ float inv_depth_at(const TermsParams& P, int d, float u, float v) {
    return fmaf(P.inv_depth[3 * d], u, fmaf(P.inv_depth[3 * d + 1], v, P.inv_depth[3 * d + 2]));
}

__device__ __forceinline__ float alpha_at(const CompositeParams& p, const float4* fb, const Taps& tp) {
    return sample_rgba(tp.kind == 2 ? fb : p.atlas_sta, tp).w;
}

// front-to-back walk over the hit planes of one ray (utils_mpi.py:100-104): blend weights bw_k = a_k * prod_{j<k}(1 - a_j)
__device__ __forceinline__ RaySums ray_sums(const TermsParams& P, const float4* fb, unsigned rem, float u, float v) {
    RaySums s{0.f, 0.f, 0.f, 0.f};
    float Tr = 1.f;
    while (rem) {
        const int d = __ffs(rem) - 1;
        rem &= rem - 1u;
        const Taps tp = make_taps(P.p, d, u, v);
        const float a = alpha_at(P.p, fb, tp);
        const float bw = a * Tr;
        s.A += bw;
        s.Dp = fmaf(bw, inv_depth_at(P, d, u, v), s.Dp);
        s.L1 += a;
        s.L2 = fmaf(a, a, s.L2);
        Tr *= 1.f - a;
    }
    return s;
}

__global__ void __launch_bounds__(BX* BY) composite_terms_fwd_kernel(const __grid_constant__ TermsParams P) {
    const CompositeParams& p = P.p;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int px = blockIdx.x * BX + tx, py = blockIdx.y * BY + ty, t = blockIdx.z;
    const int H = p.view.H, W = p.view.W;
    const bool active = px < W && py < H;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;
    const unsigned hits = active ? hit_mask(p, u, v) : 0u;
    const int ft = p.ts ? __ldg(&p.ts[t]) : t;
    const float4* fb = p.atlas_dyn + (size_t)ft * p.view.dyn_h * p.view.dyn_w;
    const RaySums s = ray_sums(P, fb, hits, u, v);
    double sp = 0.0;
    if (active) {
        const size_t o = ((size_t)t * H + py) * W + px;
        if (P.alpha_out) P.alpha_out[o] = s.A;
        if (P.disp_out) P.disp_out[o] = s.Dp;
        sp = (double)(s.L1 / fmaxf(sqrtf(s.L2), P.sparsity_eps));    // alpha.norm(p=1) / alpha.norm(p=2).clamp_min(eps)
    }
    if (P.sparsity_sum) {
        __shared__ double s_sum[(BX * BY) / 32];
        const int warp = (ty * BX + tx) >> 5;
        const double w = warp_sum(sp);
        if (tx == 0) s_sum[warp] = w;
        __syncthreads();
        if (tx == 0 && ty == 0) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < (BX * BY) / 32; ++i) acc += s_sum[i];
            atomicAdd(P.sparsity_sum, acc);
        }
    }
}

__device__ __forceinline__ void scatter_alpha(float4* gb, const Taps& tp, float gl) {
    if (tp.w00 != 0.f) atomicAdd(&gb[tp.o00].w, gl * tp.w00);
    if (tp.w10 != 0.f) atomicAdd(&gb[tp.o10].w, gl * tp.w10);
    if (tp.w01 != 0.f) atomicAdd(&gb[tp.o01].w, gl * tp.w01);
    if (tp.w11 != 0.f) atomicAdd(&gb[tp.o11].w, gl * tp.w11);
}

// With e_k = g_alpha + g_disp / z_k the upstream scalar per slot, tot = sum_k bw_k e_k and S_k = sum_{j>k} bw_j e_j:
//   d/da_k [sum_j bw_j e_j] = Tr_k e_k - S_k / (1 - a_k);   d/da_k [L1 / n] = 1/n - L1 a_k / n^3  (n = |a|_2 >= eps; else 1/eps)
// and da_k / d(logit) = a_k (1 - a_k) (sigmoid, MPV.py:435).
__global__ void __launch_bounds__(BX* BY) composite_terms_bwd_kernel(const __grid_constant__ TermsParams P) {
    const CompositeParams& p = P.p;
    const int px = blockIdx.x * BX + threadIdx.x, py = blockIdx.y * BY + threadIdx.y, t = blockIdx.z;
    const int H = p.view.H, W = p.view.W;
    if (px >= W || py >= H) return;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;
    const unsigned hits = hit_mask(p, u, v);
    if (hits == 0u) return;
    const int ft = p.ts ? __ldg(&p.ts[t]) : t;
    const size_t frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* fb = p.atlas_dyn + (size_t)ft * frame;
    float4* gdyn = p.grad_dyn + (size_t)ft * frame;
    const size_t o = ((size_t)t * H + py) * W + px;
    const float gA = P.g_alpha ? __ldg(P.g_alpha + o) : 0.f;
    const float gD = P.g_disp ? __ldg(P.g_disp + o) : 0.f;
    const float ws = P.w_sparsity ? __ldg(P.w_sparsity) : 0.f;

    // pass 1: the ray's totals, accumulated exactly as pass 2 accumulates its prefix (so the last suffix is exactly 0)
    float tot = 0.f, L1 = 0.f, L2 = 0.f;
    {
        float Tr = 1.f;
        unsigned rem = hits;
        while (rem) {
            const int d = __ffs(rem) - 1;
            rem &= rem - 1u;
            const Taps tp = make_taps(p, d, u, v);
            const float a = alpha_at(p, fb, tp);
            const float e = fmaf(gD, inv_depth_at(P, d, u, v), gA);
            tot = fmaf(a * Tr, e, tot);
            L1 += a;
            L2 = fmaf(a, a, L2);
            Tr *= 1.f - a;
        }
    }
    float c1 = 0.f, c2 = 0.f;
    if (ws != 0.f) {
        const float n = sqrtf(L2);
        if (n >= P.sparsity_eps) {
            const float inv = 1.f / n;
            c1 = ws * inv;
            c2 = ws * L1 * inv * inv * inv;
        } else {
            c1 = ws / P.sparsity_eps;                              // clamp_min: the norm gets no gradient
        }
    }
    float Tr = 1.f, pre = 0.f;
    unsigned rem = hits;
    while (rem) {
        const int d = __ffs(rem) - 1;
        rem &= rem - 1u;
        const Taps tp = make_taps(p, d, u, v);
        const float a = alpha_at(p, fb, tp);
        const float e = fmaf(gD, inv_depth_at(P, d, u, v), gA);
        pre = fmaf(a * Tr, e, pre);
        const float S = tot - pre;
        const float gs = c1 - c2 * a;
        const float gl = a * ((1.f - a) * fmaf(Tr, e, gs) - S);
        Tr *= 1.f - a;
        if (gl != 0.f) scatter_alpha(tp.kind == 2 ? gdyn : p.grad_sta, tp, gl);
    }
}

static int fill_terms(TermsParams& P, const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn, const float* atlas_sta,
                      const int32_t* ts, int32_t T, const float* inv_depth_host, float sparsity_eps) {
    if (int e = validate_view(view, quads, atlas_dyn, atlas_sta)) return e;
    VL3D_REQUIRE(T >= 1, VL3D_EINVAL, "bad T=%d", T);
    VL3D_REQUIRE(sparsity_eps > 0.f, VL3D_EINVAL, "sparsity_eps must be positive");
    CompositeParams& p = P.p;
    p.view = *view; p.quads = quads;
    p.atlas_dyn = reinterpret_cast<const float4*>(atlas_dyn);
    p.atlas_sta = reinterpret_cast<const float4*>(atlas_sta);
    p.ts = ts; p.T = T; p.pad = 0;
    for (int i = 0; i < VL3D_MAX_PLANES * 3; ++i) P.inv_depth[i] = (inv_depth_host && i < 3 * view->D) ? inv_depth_host[i] : 0.f;
    P.sparsity_eps = sparsity_eps;
    return 0;
}

}  // namespace vl3d

using namespace vl3d;

extern "C" int vl3d_composite_terms_fwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                                        const float* atlas_sta, const int32_t* ts, int32_t T, const float* inv_depth_host,
                                        float sparsity_eps, float* alpha_out, float* disp_out, double* sparsity_sum,
                                        void* stream) {
    TermsParams P{};
    if (int e = fill_terms(P, view, quads, atlas_dyn, atlas_sta, ts, T, inv_depth_host, sparsity_eps)) return e;
    VL3D_REQUIRE(alpha_out || disp_out || sparsity_sum, VL3D_ENULL, "composite_terms_fwd: no output requested");
    VL3D_REQUIRE(disp_out == nullptr || inv_depth_host != nullptr, VL3D_ENULL, "composite_terms_fwd: disp_out needs inv_depth_host");
    P.alpha_out = alpha_out; P.disp_out = disp_out; P.sparsity_sum = sparsity_sum;
    dim3 grid((view->W + BX - 1) / BX, (view->H + BY - 1) / BY, T), block(BX, BY);
    composite_terms_fwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P);
    return check_launch("composite_terms_fwd");
}

extern "C" int vl3d_composite_terms_bwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                                        const float* atlas_sta, const int32_t* ts, int32_t T, const float* inv_depth_host,
                                        float sparsity_eps, const float* grad_alpha, const float* grad_disp,
                                        const float* w_sparsity, float* grad_dyn, float* grad_sta, void* stream) {
    TermsParams P{};
    if (int e = fill_terms(P, view, quads, atlas_dyn, atlas_sta, ts, T, inv_depth_host, sparsity_eps)) return e;
    VL3D_REQUIRE(grad_dyn != nullptr || atlas_dyn == nullptr, VL3D_ENULL, "grad_dyn is NULL");
    VL3D_REQUIRE(grad_sta != nullptr || atlas_sta == nullptr, VL3D_ENULL, "grad_sta is NULL");
    VL3D_REQUIRE(((uintptr_t)grad_dyn & 15) == 0 && ((uintptr_t)grad_sta & 15) == 0, VL3D_EALIGN,
                 "gradient pointers must be 16-byte aligned");
    VL3D_REQUIRE(grad_disp == nullptr || inv_depth_host != nullptr, VL3D_ENULL, "composite_terms_bwd: grad_disp needs inv_depth_host");
    if (!grad_alpha && !grad_disp && !w_sparsity) return 0;         // nothing upstream
    P.p.grad_dyn = reinterpret_cast<float4*>(grad_dyn); P.p.grad_sta = reinterpret_cast<float4*>(grad_sta);
    P.g_alpha = grad_alpha; P.g_disp = grad_disp; P.w_sparsity = w_sparsity;
    dim3 grid((view->W + BX - 1) / BX, (view->H + BY - 1) / BY, T), block(BX, BY);
    composite_terms_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P);
    return check_launch("composite_terms_bwd");
}
