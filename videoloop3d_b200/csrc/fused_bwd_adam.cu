// Fused composite backward + Adam for the dynamic atlas: ONE persistent kernel per optimisation step.
//
// Replaces (reference file:line): autograd's backward through MPV.py:413-451 / utils_mpi.py:92-107 / MPV.py:517-531
// followed by `optimizer.zero_grad()` / `optimizer.step()` of train_3dvid.py:242-244 with
// torch.optim.Adam(betas=(0.9,0.999), eps=6e-8) (MPV.py:200-218)  — SURVEY.md §8(f) N4.
//
// Why one kernel: as separate launches the texel gradient crosses HBM three times (zero-fill 22.6 GB written,
// RED read-modify-write 43 GB, Adam read 22.6 GB at 720p / D=32 / T=48) and an issue-bound kernel (backward: 2.4 TB/s)
// alternates with a DRAM-bound one (Adam: 6.3 TB/s).  Here resident CTAs pull work items from one ordered queue:
//   BWD   one (screen tile, chunk of TF frames) of the backward — exactly composite_lean.cuh's bwd_tile;
//   ADAM  Adam on a rectangle of texels (rows x width) of the chunk's frames: reads p, g, m, v; writes p, m, v; then
//         writes zeros back into g unless the schedule zeroes ahead (then the buffer's content between steps does not
//         matter).  (Dropping the consumed lines with discard.global.L2 was measured 3x slower on B200: not used.)
//   ZERO  zero a rectangle of g shortly before the first tile that accumulates into it (full-line stores: the RED
//         that follows hits L2 instead of fetching the line from DRAM).
// so backward tiles (instruction issue) and Adam rectangles (DRAM) overlap on every SM, and — when the host orders the
// items by screen band (dense layout) — a band's gradient rows are produced, consumed and dropped while L2-resident.
//
// The schedule is a host-built table (videoloop3d_b200/schedule.py): per item the work description, up to two ranges
// of counters to wait for (each >= a target) and a counter to bump when done.  An item only waits for items that precede
// it in the queue, and items are claimed in queue order by running CTAs, so the waits cannot deadlock.  One table
// describes one "round" (= one chunk of TF frames); the kernel replays it for every chunk, with per-chunk counters.
#include "composite_common.cuh"
#include "tma_common.cuh"
#include "composite_lean.cuh"

namespace vl3d {

constexpr int FUSED_TF = 2;

enum : int { ITEM_BWD = 0, ITEM_ADAM = 1, ITEM_ZERO = 2 };
enum : int { FLAG_HAS_GRAD = 1, FLAG_REZERO = 2, FLAG_PREV_ROUND = 8 };

struct alignas(64) FusedParams {
    TmaRenderParams R;
    const int4* items;     // [n_items][3]
    int n_items, n_rounds, n_chunks, n_counters;
    int* counters;         // [n_chunks][n_counters], initialised by the caller
    int* ticket;           // queue head, zeroed by the caller
    float4* m;
    float4* v;
    AdamK k;
    OwnParams own;         // owner mode only (see composite_lean.cuh)
};

__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_inc(int* p) {
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}

// Adam on rows x width texels starting at texel `base` of frames [t0, t0 + FUSED_TF) (row stride = dyn_w).
// Every thread loads ADAM_U texels x FUSED_TF frames x (p, m, v, g) before it computes.  Where the Adam items set the pace
// (tile-culled models: few samples per tile; or the dense layout while it was pinned at the DRAM limit with the plain
// gradient buffer) 16 loads in flight per thread beat 8: 17.8 vs 18.6 ms for the sparse bench model.  For the dense
// layout with the compressible gradient buffer (212 GB, DRAM at 65 %) the lighter items win: 40.9 vs 41.9 ms at 720p,
// 3.04 vs 3.13 ms at 180x320.
template <int MODE>
struct AdamLoads { static constexpr int U = MODE >= 2 ? 1 : 2; };

// Adam on an explicit list of texels (offsets from `base`, the same for every frame of the chunk): ADAM_U texels x
// FUSED_TF frames x (p, m, v, g) loads in flight per thread, like the dense loop below.
template <int ADAM_U>
__device__ __forceinline__ void adam_list(const FusedParams& F, float4* const P0, float4* const G0, float4* const M0, float4* const V0,
                                          const size_t frame, const int* list, const int n, const bool has_grad, const bool rezero) {
    const int tid = threadIdx.y * BX + threadIdx.x;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int NT = BX * BY, NV = ADAM_U * FUSED_TF;
    for (int q0 = tid; q0 < n; q0 += NT * ADAM_U) {
        float4 pp[NV], gg[NV], mm[NV], vv[NV];
        size_t off[NV];
        bool in[NV];
#pragma unroll
        for (int u = 0; u < ADAM_U; ++u) {
            const int q = q0 + u * NT;
            const bool todo = q < n;
            const int o = todo ? list[q] : 0;
#pragma unroll
            for (int f = 0; f < FUSED_TF; ++f) {
                const int k = u * FUSED_TF + f;
                in[k] = todo;
                off[k] = (size_t)f * frame + o;
                pp[k] = mm[k] = vv[k] = gg[k] = zero4;
                if (todo) {
                    pp[k] = __ldcs(P0 + off[k]);
                    mm[k] = __ldcs(M0 + off[k]);
                    vv[k] = __ldcs(V0 + off[k]);
                    if (has_grad) gg[k] = __ldcg(G0 + off[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (in[k]) {
                adam4(pp[k], gg[k], mm[k], vv[k], F.k);
                __stcs(P0 + off[k], pp[k]);
                __stcs(M0 + off[k], mm[k]);
                __stcs(V0 + off[k], vv[k]);
                if (rezero) G0[off[k]] = zero4;
            }
        }
    }
}

// OWN: `plane` >= 0 says the rectangle lies in that plane's atlas cell; texels a screen tile owns were updated by the tile
// (bwd_tile's own_flush) and are skipped here.  Only ~1/3 of the texels are left, two or three per tile and atlas row
// plus whole rows between the tile rows: the CTA first compacts the texels it has to do (OWN_LIST positions at a time,
// classification only) and then runs Adam over the list with every lane busy.
constexpr int OWN_LIST = 2048;

template <bool OWN, int ADAM_U>
__device__ __forceinline__ void adam_rect(const FusedParams& F, const int t0, const int base, const int width, const int rows,
                                          const int flags, const int plane) {
    const CompositeParams& p = F.R.p;
    const int tid = threadIdx.y * BX + threadIdx.x;
    const size_t frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const int stride = p.view.dyn_w;
    float4* const P0 = const_cast<float4*>(p.atlas_dyn) + (size_t)t0 * frame + base;
    float4* const G0 = p.grad_dyn + (size_t)t0 * frame + base;
    float4* const M0 = F.m + (size_t)t0 * frame + base;
    float4* const V0 = F.v + (size_t)t0 * frame + base;
    const bool has_grad = (flags & FLAG_HAS_GRAD) != 0;
    const bool rezero = has_grad && (flags & FLAG_REZERO);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int NT = BX * BY, NV = ADAM_U * FUSED_TF;
    if (OWN && plane >= 0) {
        __shared__ int s_list[OWN ? OWN_LIST : 1];
        __shared__ int s_n;
        const int y0 = base / stride, x0 = base - y0 * stride;
        const int n = rows * width;
        const int lane = tid & 31;
        for (int b0 = 0; b0 < n; b0 += OWN_LIST) {
            if (tid == 0) s_n = 0;
            __syncthreads();
#pragma unroll 2
            for (int i = b0 + tid; i < min(b0 + OWN_LIST, n); i += NT) {   // (the trip count is warp-uniform up to the tail)
                const int r = i / width, c = i - r * width;
                const bool keep = !own_texel_of_any(F.own, plane, x0 + c, y0 + r);
                const unsigned act = __activemask();
                const unsigned m = __ballot_sync(act, keep);
                int pos = 0;
                if (lane == __ffs(act) - 1) pos = atomicAdd(&s_n, __popc(m));
                pos = __shfl_sync(act, pos, __ffs(act) - 1);
                if (keep) s_list[pos + __popc(m & ((1u << lane) - 1u))] = r * stride + c;
            }
            __syncthreads();
            adam_list<ADAM_U>(F, P0, G0, M0, V0, frame, s_list, s_n, has_grad, rezero);
            __syncthreads();
        }
        return;
    }
    for (int r = 0; r < rows; ++r) {
        const size_t ro = (size_t)r * stride;
        for (int c0 = tid; c0 < width; c0 += NT * ADAM_U) {
            float4 pp[NV], gg[NV], mm[NV], vv[NV];
            size_t off[NV];
            bool in[NV];
#pragma unroll
            for (int u = 0; u < ADAM_U; ++u)
#pragma unroll
                for (int f = 0; f < FUSED_TF; ++f) {
                    const int c = c0 + u * NT, k = u * FUSED_TF + f;
                    in[k] = c < width;
                    off[k] = (size_t)f * frame + ro + c;
                    pp[k] = mm[k] = vv[k] = gg[k] = zero4;
                    if (in[k]) {                                    // streams: evict-first; the gradient: L2 only
                        pp[k] = __ldcs(P0 + off[k]);
                        mm[k] = __ldcs(M0 + off[k]);
                        vv[k] = __ldcs(V0 + off[k]);
                        if (has_grad) gg[k] = __ldcg(G0 + off[k]);
                    }
                }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (in[k]) {
                    adam4(pp[k], gg[k], mm[k], vv[k], F.k);
                    __stcs(P0 + off[k], pp[k]);
                    __stcs(M0 + off[k], mm[k]);
                    __stcs(V0 + off[k], vv[k]);
                    if (rezero) G0[off[k]] = zero4;
                }
            }
        }
    }
}

__device__ __forceinline__ void zero_rect(const FusedParams& F, const int t0, const int base, const int width, const int rows) {
    const CompositeParams& p = F.R.p;
    const int tid = threadIdx.y * BX + threadIdx.x;
    const size_t frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const int stride = p.view.dyn_w;
    float4* const G0 = p.grad_dyn + (size_t)t0 * frame + base;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rows; ++r)
        for (int c = tid; c < width; c += BX * BY) {
#pragma unroll
            for (int f = 0; f < FUSED_TF; ++f) G0[(size_t)f * frame + (size_t)r * stride + c] = zero4;
        }
}

// one warp per screen tile (regulariser tiling): zone[tile][plane] = where the tile owns texels of the plane (x = -1: it
// does not), table[tile] = the planes it owns.  The only place where ownership of a (tile, plane) pair is decided.
static __global__ void own_table_kernel(const __grid_constant__ FusedParams F, unsigned* __restrict__ table, int2* __restrict__ zone) {
    const CompositeParams& p = F.R.p;
    const int gx = F.own.gx, gy = F.own.gy;
    const int tile = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (tile >= gx * gy) return;
    const int by = tile / gx, bx = tile - by * gx;
    int cls = 0;
    int4 box = make_int4(0, 0, 0, 0);
    if (lane < p.view.D)
        cls = tile_plane_class(p, lane, bx * (BX - 1), min(bx * (BX - 1) + BX - 1, p.view.W - 1), by * (BY - 1),
                               min(by * (BY - 1) + BY - 1, p.view.H - 1), box);
    const unsigned m_mixed = __ballot_sync(0xffffffffu, cls == 2);
    int2 z = make_int2(-1, -1);
    if (lane < p.view.D && m_mixed == 0u) z = tile_own_zone(p, F.own, lane, bx, by, cls, box);
    zone[(size_t)tile * VL3D_MAX_PLANES + lane] = z;
    const unsigned m_own = __ballot_sync(0xffffffffu, z.x >= 0);
    if (lane == 0) table[tile] = m_own;
}

template <bool SMOOTH, int MODE, bool OWN>
__global__ void __launch_bounds__(BX* BY, 3) fused_bwd_adam_kernel(const __grid_constant__ FusedParams F) {
    __shared__ int s_item;
    const int tid = threadIdx.y * BX + threadIdx.x;
    const long long total = (long long)F.n_rounds * F.n_items;
    unsigned kbase = 0u;
    bool first = true;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(F.ticket, 1);
        __syncthreads();
        const int I = s_item;
        __syncthreads();
        if (I >= total) break;
        const int round = I / F.n_items, j = I - round * F.n_items;
        const int4 a = __ldg(&F.items[3 * j]), b = __ldg(&F.items[3 * j + 1]), c = __ldg(&F.items[3 * j + 2]);
        const int type = a.x & 15, flags = a.x >> 4;
        const int chunk = round - ((flags & FLAG_PREV_ROUND) ? 1 : 0);
        if (chunk < 0 || chunk >= F.n_chunks) continue;
        int* const cnt = F.counters + (size_t)chunk * F.n_counters;
        if (b.y > 0 || c.y > 0) {                                   // wait: counters [b.x, b.x + b.y) >= b.z, [c.x, c.x + c.y) >= c.z
            if (tid < 32) {
                unsigned ns = 32;
                for (;;) {
                    bool ok = true;
                    for (int i = tid; i < b.y; i += 32) ok = ok && (ld_relaxed_gpu(cnt + b.x + i) >= b.z);
                    for (int i = tid; i < c.y; i += 32) ok = ok && (ld_relaxed_gpu(cnt + c.x + i) >= c.z);
                    if (__all_sync(0xffffffffu, ok)) break;
                    __nanosleep(ns);
                    if (ns < 1024) ns *= 2;
                }
                __threadfence();                                    // acquire side of the producers' release increments
            }
            __syncthreads();
        }
        const int t0 = chunk * FUSED_TF;
        if (type == ITEM_BWD) {
            bwd_tile<FUSED_TF, SMOOTH, MODE, OWN>(F.R, a.y, a.z, t0, kbase, first, &F.own);
            first = false;
        } else if (type == ITEM_ADAM) {
            adam_rect<OWN, AdamLoads<MODE>::U>(F, t0, a.y, a.z, a.w, flags, c.w);
        } else {
            zero_rect(F, t0, a.y, a.z, a.w);
        }
        if (b.w >= 0) {                                             // signal: this item's writes / REDs are visible
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                red_release_gpu_inc(cnt + b.w);
            }
        }
    }
}

// bytes of scratch one CTA of the owner-mode kernel needs, and how many CTAs a launch may use (occupancy x SMs)
static size_t own_scratch_per_cta() { return (size_t)2 * FUSED_TF * OWN_ZN * sizeof(float4); }

template <bool SMOOTH, int MODE, bool OWN>
static int launch_fused(const FusedParams& F, int ctas_per_sm, cudaStream_t st, size_t scratch_bytes = 0) {
    const size_t smem = MODE >= 2 ? (size_t)BWD_TMA_STAGES * FUSED_TF * TMA_BOX_BYTES : 0;
    auto kern = fused_bwd_adam_kernel<SMOOTH, MODE, OWN>;
    if (smem) {
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return set_err((int)ce, "fused_bwd_adam: cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    }
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BX * BY, smem);
    if (occ < 1) return set_err(VL3D_EINVAL, "fused_bwd_adam: kernel does not fit on an SM");
    if (ctas_per_sm >= 1 && ctas_per_sm < occ) occ = ctas_per_sm;
    long long grid = (long long)sms * occ;
    const long long total = (long long)F.n_rounds * F.n_items;
    if (grid > total) grid = total;
    if (OWN) {
        if ((size_t)grid * own_scratch_per_cta() > scratch_bytes)
            return set_err(VL3D_EINVAL, "fused_bwd_adam_own: scratch of %zu bytes is too small for %lld CTAs x %zu bytes", scratch_bytes,
                           grid, own_scratch_per_cta());
        const int tiles = F.own.gx * F.own.gy;
        own_table_kernel<<<(tiles * 32 + 255) / 256, 256, 0, st>>>(F, const_cast<unsigned*>(F.own.table), const_cast<int2*>(F.own.zone));
        if (int e = check_launch("own_table")) return e;
    }
    kern<<<(unsigned)grid, dim3(BX, BY), smem, st>>>(F);
    return check_launch("fused_bwd_adam");
}

}  // namespace vl3d

using namespace vl3d;

extern "C" int64_t vl3d_fused_own_table_bytes(int32_t H, int32_t W) {
    const int64_t tiles = (int64_t)((W + BX - 2) / (BX - 1)) * ((H + BY - 2) / (BY - 1));
    return 4 * (((tiles + 3) & ~(int64_t)3) + tiles * VL3D_MAX_PLANES * 2);
}

static int fused_entry(const vl3d_view* view, const vl3d_quad* quads, float* atlas_dyn, const float* atlas_sta, int32_t T,
                       const float* grad_rgb, const float* rgb, const float* w_smooth, double* smooth_sums, float* grad_dyn,
                       float* grad_sta, float* adam_m, float* adam_v, int32_t step, float lr, float beta1, float beta2, float eps,
                       const int32_t* items, int32_t n_items, int32_t n_rounds, int32_t* counters, int32_t n_counters,
                       int32_t* ticket, int32_t ctas_per_sm, const vl3d_own* own, uint32_t* own_table, int64_t own_table_bytes,
                       float* scratch, int64_t scratch_bytes, void* stream) {
    if (int e = validate_view(view, quads, atlas_dyn, atlas_sta)) return e;
    VL3D_REQUIRE(grad_rgb && rgb && grad_dyn && adam_m && adam_v && atlas_dyn, VL3D_ENULL, "fused_bwd_adam: NULL pointer");
    VL3D_REQUIRE(grad_sta != nullptr || atlas_sta == nullptr, VL3D_ENULL, "grad_sta is NULL");
    VL3D_REQUIRE(items && counters && ticket, VL3D_ENULL, "fused_bwd_adam: schedule pointers are NULL");
    VL3D_REQUIRE(T >= FUSED_TF && T % FUSED_TF == 0, VL3D_EINVAL, "fused_bwd_adam: T=%d must be a positive multiple of %d", T, FUSED_TF);
    VL3D_REQUIRE(n_items >= 1 && n_counters >= 1 && n_rounds >= T / FUSED_TF && n_rounds <= T / FUSED_TF + 1, VL3D_EINVAL,
                 "fused_bwd_adam: bad schedule (n_items=%d n_rounds=%d n_counters=%d)", n_items, n_rounds, n_counters);
    VL3D_REQUIRE(step >= 1, VL3D_EINVAL, "fused_bwd_adam: step=%d", step);
    VL3D_REQUIRE((((uintptr_t)grad_dyn | (uintptr_t)grad_sta | (uintptr_t)adam_m | (uintptr_t)adam_v | (uintptr_t)items) & 15) == 0,
                 VL3D_EALIGN, "fused_bwd_adam: pointers must be 16-byte aligned");
    FusedParams F{};
    CompositeParams& p = F.R.p;
    p.view = *view; p.quads = quads;
    p.atlas_dyn = reinterpret_cast<const float4*>(atlas_dyn);
    p.atlas_sta = reinterpret_cast<const float4*>(atlas_sta);
    p.ts = nullptr; p.T = T; p.pad = 0; p.tb = 0;
    p.grad_rgb = grad_rgb; p.rgb = rgb;
    p.grad_dyn = reinterpret_cast<float4*>(grad_dyn); p.grad_sta = reinterpret_cast<float4*>(grad_sta);
    p.w_smooth = w_smooth; p.smooth = smooth_sums;
    VL3D_REQUIRE(w_smooth != nullptr || smooth_sums == nullptr, VL3D_EINVAL, "smooth_sums needs w_smooth");
    F.items = reinterpret_cast<const int4*>(items);
    F.n_items = n_items; F.n_rounds = n_rounds; F.n_chunks = T / FUSED_TF; F.n_counters = n_counters;
    F.counters = counters; F.ticket = ticket;
    F.m = reinterpret_cast<float4*>(adam_m); F.v = reinterpret_cast<float4*>(adam_v);
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    F.k.b1 = beta1; F.k.b2 = beta2; F.k.eps = eps;
    F.k.step_size = (float)((double)lr / bc1);
    F.k.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    cudaStream_t st = (cudaStream_t)stream;
    const bool smooth = w_smooth != nullptr;
    if (own != nullptr) {
        VL3D_REQUIRE(smooth && (view->flags & VL3D_VIEW_RECT_PLANES), VL3D_EINVAL,
                     "fused_bwd_adam_own needs the dense layout (VL3D_VIEW_RECT_PLANES) and the regulariser weights");
        VL3D_REQUIRE(own_table && scratch, VL3D_ENULL, "fused_bwd_adam_own: own_table / scratch is NULL");
        VL3D_REQUIRE((((uintptr_t)scratch | (uintptr_t)own_table) & 15) == 0, VL3D_EALIGN,
                     "fused_bwd_adam_own: scratch / own_table must be 16-byte aligned");
        VL3D_REQUIRE(own_table_bytes >= vl3d_fused_own_table_bytes(view->H, view->W), VL3D_EINVAL,
                     "fused_bwd_adam_own: own_table of %lld bytes is too small (%lld needed)", (long long)own_table_bytes,
                     (long long)vl3d_fused_own_table_bytes(view->H, view->W));
        VL3D_REQUIRE(make_atlas_tmap(&F.R.tmap, p.view, atlas_dyn, T), VL3D_EINVAL, "fused_bwd_adam_own: no tensor map for the atlas");
        OwnParams& O = F.own;
        for (int i = 0; i < VL3D_MAX_PLANES * 9; ++i) O.hinv[i] = own->hinv[i];
        for (int d = 0; d < VL3D_MAX_PLANES; ++d) {
            O.L[d] = own->reach[d];
            O.rect[d] = make_int4(own->rect[4 * d], own->rect[4 * d + 1], own->rect[4 * d + 2], own->rect[4 * d + 3]);
        }
        O.scratch = reinterpret_cast<float4*>(scratch);
        O.gx = (view->W + BX - 2) / (BX - 1); O.gy = (view->H + BY - 2) / (BY - 1);
        O.table = own_table;
        O.zone = reinterpret_cast<const int2*>(own_table + (((size_t)O.gx * O.gy + 3) & ~(size_t)3));
        O.m = F.m; O.v = F.v; O.k = F.k;
        return launch_fused<true, 3, true>(F, ctas_per_sm, st, (size_t)(scratch_bytes < 0 ? 0 : scratch_bytes));
    }
    if (smooth && (view->flags & VL3D_VIEW_RECT_PLANES) && make_atlas_tmap(&F.R.tmap, p.view, atlas_dyn, T))
        return launch_fused<true, 3, false>(F, ctas_per_sm, st);
    if (smooth) return launch_fused<true, 0, false>(F, ctas_per_sm, st);
    return launch_fused<false, 0, false>(F, ctas_per_sm, st);
}

extern "C" int vl3d_fused_bwd_adam(const vl3d_view* view, const vl3d_quad* quads, float* atlas_dyn, const float* atlas_sta,
                                   int32_t T, const float* grad_rgb, const float* rgb, const float* w_smooth,
                                   double* smooth_sums, float* grad_dyn, float* grad_sta, float* adam_m, float* adam_v,
                                   int32_t step, float lr, float beta1, float beta2, float eps, const int32_t* items,
                                   int32_t n_items, int32_t n_rounds, int32_t* counters, int32_t n_counters,
                                   int32_t* ticket, int32_t ctas_per_sm, void* stream) {
    return fused_entry(view, quads, atlas_dyn, atlas_sta, T, grad_rgb, rgb, w_smooth, smooth_sums, grad_dyn, grad_sta, adam_m,
                       adam_v, step, lr, beta1, beta2, eps, items, n_items, n_rounds, counters, n_counters, ticket, ctas_per_sm,
                       nullptr, nullptr, 0, nullptr, 0, stream);
}

extern "C" int64_t vl3d_fused_own_scratch_bytes(int32_t ctas) { return (int64_t)ctas * (int64_t)own_scratch_per_cta(); }

extern "C" int vl3d_fused_bwd_adam_own(const vl3d_view* view, const vl3d_quad* quads, float* atlas_dyn, const float* atlas_sta,
                                       int32_t T, const float* grad_rgb, const float* rgb, const float* w_smooth,
                                       double* smooth_sums, float* grad_dyn, float* grad_sta, float* adam_m, float* adam_v,
                                       int32_t step, float lr, float beta1, float beta2, float eps, const int32_t* items,
                                       int32_t n_items, int32_t n_rounds, int32_t* counters, int32_t n_counters,
                                       int32_t* ticket, int32_t ctas_per_sm, const vl3d_own* own, uint32_t* own_table,
                                       int64_t own_table_bytes, float* scratch, int64_t scratch_bytes, void* stream) {
    VL3D_REQUIRE(own != nullptr, VL3D_ENULL, "fused_bwd_adam_own: own is NULL");
    return fused_entry(view, quads, atlas_dyn, atlas_sta, T, grad_rgb, rgb, w_smooth, smooth_sums, grad_dyn, grad_sta, adam_m,
                       adam_v, step, lr, beta1, beta2, eps, items, n_items, n_rounds, counters, n_counters, ticket, ctas_per_sm,
                       own, own_table, own_table_bytes, scratch, scratch_bytes, stream);
}
