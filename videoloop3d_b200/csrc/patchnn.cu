// Looping loss for sm_100a: temporal patch nearest-neighbour search, vote/merge, robust loss.
//
// Replaces (reference file:line): utils_vid.py:60-69 extract_3Dpatches (unfoldNd im2col),
// :72-86 efficient_compute_distances (bmm), :109-119 get_col_mins_efficient, :122-142
// get_NN_indices_low_memory, :206-229 FindNNpatchAndMerge (gather + FoldNd), :289-349 the macro-block
// loop of Patch3DGPNNLowMemLoss (a memory trick that does not change results), :10-26 robust_lossfun,
// and MPV.py:499-504 (scale-invariant gain).
//
// Search: the unfold -> GEMM pipeline is, per spatial patch position, a frame-pair distance matrix
// G[tx][ty] = sum over the p x p x 3 window of (x[tx] - y[ty])^2 followed by a pt-long diagonal sum
// D[i][j] = sum_dt G[i*st+dt][j*st+dt].  One CTA owns one patch position; it streams the window one
// pixel row (3 channels) at a time through shared memory, every thread keeps a 4x4 (tx,ty) register
// tile of G, and candidates are consumed in chunks of 64 frames so the GPNN column-min normaliser
// (min over all queries) and the running first-min argmin never need the whole matrix.
// Distances are direct fp32 sums of squared differences (no |x|^2+|y|^2-2xy cancellation, no
// tensor cores): index parity with an fp64 search is limited only by exact near-ties.
#include <stdlib.h>

#include "vl3d_common.cuh"
#include "tma_ptx.cuh"

namespace vl3d {

constexpr int NN_THREADS = 256;
constexpr int NN_CF = 64;      // frames per chunk (both x block and y chunk)
constexpr int NN_R = 4;        // register tile edge
constexpr int NN_MAX_N1 = 256;
constexpr int NB = 3;         // strip kernel: staging buffers (rows in flight = NB - 1)

struct SearchParams {
    vl3d_loss_desc d;
    const float* x;
    const float* y;
    int* nn;
    int groups;     // float4 groups per slab = ceil(3p/4)
    int row0;       // first patch row handled by this launch
};

// smem carve-up (floats): xs4[2][groups][64][4], ys4[2][groups][64][4], Gs[64][65], Ds[n1][64],
// colmin[64], best_val[n1], best_idx[n1]
__global__ void __launch_bounds__(NN_THREADS) patchnn_search_kernel(const __grid_constant__ SearchParams P) {
    extern __shared__ __align__(16) float smem[];
    const vl3d_loss_desc& L = P.d;
    const int G4 = P.groups;
    float4* xs4 = reinterpret_cast<float4*>(smem);                  // [2][G4][64]
    float4* ys4 = xs4 + 2 * G4 * NN_CF;                             // [2][G4][64]
    float* Gs = reinterpret_cast<float*>(ys4 + 2 * G4 * NN_CF);     // [64][65]
    float* Ds = Gs + NN_CF * (NN_CF + 1);                           // [n1][64]
    float* colmin = Ds + (size_t)L.n1 * NN_CF;                      // [64]
    float* best_val = colmin + NN_CF;                               // [n1]
    int* best_idx = reinterpret_cast<int*>(best_val + L.n1);        // [n1]

    const int tid = threadIdx.x;
    const int pxi = blockIdx.x, pyi = P.row0 + blockIdx.y;                   // patch position on the stride-s grid
    const int x0 = pxi * L.s, y0 = pyi * L.s;
    const int p = L.p, pt = L.pt, st = L.st;
    const float inv_d = 1.f / (float)(3 * pt * p * p);              // dist /= d (utils_vid.py:83-84)
    const int tb = tid & 15, ta = tid >> 4;                         // thread tile: x frames ta+16*i, y frames tb+16*j

    for (int i = tid; i < L.n1; i += NN_THREADS) { best_val[i] = INFINITY; best_idx[i] = 0; }

    const int tx_used = (L.n1 - 1) * st + pt, ty_used = (L.n2 - 1) * st + pt;
    // candidate chunks: y frames [c0, c0+64); candidates j in [j0, j1) fully inside the chunk
    for (int j0 = 0; j0 < L.n2;) {
        const int c0 = j0 * st;
        int j1 = (c0 + NN_CF - pt) / st + 1;                        // last j with j*st+pt-1 <= c0+63, exclusive
        if (j1 > L.n2) j1 = L.n2;
        const int cj = j1 - j0;
        // query blocks: x frames [a0, a0+64)
        for (int i0 = 0; i0 < L.n1;) {
            const int a0 = i0 * st;
            int i1 = (a0 + NN_CF - pt) / st + 1;
            if (i1 > L.n1) i1 = L.n1;

            float acc[NN_R][NN_R];
#pragma unroll
            for (int i = 0; i < NN_R; ++i)
#pragma unroll
                for (int j = 0; j < NN_R; ++j) acc[i][j] = 0.f;

            // stage one slab (pixel row `row` of the window, 3 channels) into buffer `buf`
            auto stage = [&](int row, int buf) {
                // rows of p contiguous pixels: id -> (which, channel, frame)
                for (int id = tid; id < 2 * 3 * NN_CF; id += NN_THREADS) {
                    const int fr = id & (NN_CF - 1);
                    const int c = (id >> 6) % 3;
                    const int which = id / (3 * NN_CF);               // 0: x, 1: y
                    float* dst = reinterpret_cast<float*>((which ? ys4 : xs4) + (size_t)buf * G4 * NN_CF);
                    const int gf = (which ? c0 : a0) + fr;            // global frame
                    const bool ok = gf < (which ? ty_used : tx_used);
                    const float* src = which ? P.y + (size_t)gf * L.y_sf + (size_t)c * L.y_sc + (size_t)(y0 + row) * L.y_sr + x0
                                             : P.x + (size_t)gf * L.x_sf + (size_t)c * L.x_sc + (size_t)(y0 + row) * L.x_sr + x0;
                    for (int dx = 0; dx < p; ++dx) {
                        const int e = c * p + dx;
                        const float val = ok ? __ldg(src + dx) : 0.f;
                        dst[((e >> 2) * NN_CF + fr) * 4 + (e & 3)] = val;
                    }
                }
                // zero the padding lanes of the last float4 group (3p..4*G4)
                const int npad = 4 * G4 - 3 * p;
                for (int id = tid; id < 2 * npad * NN_CF; id += NN_THREADS) {
                    const int fr = id & (NN_CF - 1);
                    const int e = 3 * p + (id >> 6) % npad;
                    const int which = id / (npad * NN_CF);
                    float* dst = reinterpret_cast<float*>((which ? ys4 : xs4) + (size_t)buf * G4 * NN_CF);
                    dst[((e >> 2) * NN_CF + fr) * 4 + (e & 3)] = 0.f;
                }
            };

            __syncthreads();            // previous users of the staging buffers / Gs are done
            stage(0, 0);
            __syncthreads();
            for (int row = 0; row < p; ++row) {
                const int buf = row & 1;
                if (row + 1 < p) stage(row + 1, buf ^ 1);
                const float4* xb = xs4 + (size_t)buf * G4 * NN_CF;
                const float4* yb = ys4 + (size_t)buf * G4 * NN_CF;
                for (int g = 0; g < G4; ++g) {
                    float4 xa[NN_R], ya[NN_R];
#pragma unroll
                    for (int i = 0; i < NN_R; ++i) xa[i] = xb[g * NN_CF + ta + 16 * i];
#pragma unroll
                    for (int j = 0; j < NN_R; ++j) ya[j] = yb[g * NN_CF + tb + 16 * j];
#pragma unroll
                    for (int i = 0; i < NN_R; ++i)
#pragma unroll
                        for (int j = 0; j < NN_R; ++j) {
                            float dlt;
                            dlt = xa[i].x - ya[j].x; acc[i][j] = fmaf(dlt, dlt, acc[i][j]);
                            dlt = xa[i].y - ya[j].y; acc[i][j] = fmaf(dlt, dlt, acc[i][j]);
                            dlt = xa[i].z - ya[j].z; acc[i][j] = fmaf(dlt, dlt, acc[i][j]);
                            dlt = xa[i].w - ya[j].w; acc[i][j] = fmaf(dlt, dlt, acc[i][j]);
                        }
                }
                __syncthreads();
            }
            // G tile -> shared
#pragma unroll
            for (int i = 0; i < NN_R; ++i)
#pragma unroll
                for (int j = 0; j < NN_R; ++j) Gs[(ta + 16 * i) * (NN_CF + 1) + tb + 16 * j] = acc[i][j];
            __syncthreads();
            // D[i][j] = sum_dt G[i*st+dt-a0][j*st+dt-c0] / d      (3-D patch = pt stacked 2-D windows)
            const int ni = i1 - i0;
            for (int id = tid; id < ni * cj; id += NN_THREADS) {
                const int il = id / cj, jl = id - il * cj;
                const int gx = (i0 + il) * st - a0, gy = (j0 + jl) * st - c0;
                float sum = 0.f;
                for (int dt = 0; dt < pt; ++dt) sum += Gs[(gx + dt) * (NN_CF + 1) + gy + dt];
                Ds[(size_t)(i0 + il) * NN_CF + jl] = sum * inv_d;
            }
            i0 = i1;
        }
        __syncthreads();
        // GPNN normaliser: D[i][j] / (alpha + min_i D[i][j])   (utils_vid.py:118,133-141)
        if (L.use_alpha) {
            for (int jl = tid; jl < cj; jl += NN_THREADS) {
                float m = INFINITY;
                for (int i = 0; i < L.n1; ++i) {
                    const float vv = Ds[(size_t)i * NN_CF + jl];
                    m = (vv < m || vv != vv) ? vv : m;               // NaN propagates like torch.min
                }
                colmin[jl] = L.alpha + m;
            }
            __syncthreads();
        }
        // running first-min argmin over candidates (torch.argmin: first minimal index; NaN counts as min)
        for (int i = tid; i < L.n1; i += NN_THREADS) {
            float bv = best_val[i];
            int bi = best_idx[i];
            for (int jl = 0; jl < cj; ++jl) {
                float vv = Ds[(size_t)i * NN_CF + jl];
                if (L.use_alpha) vv = vv / colmin[jl];
                const bool better = (vv < bv) || (vv != vv && bv == bv);
                if (better) { bv = vv; bi = j0 + jl; }
            }
            best_val[i] = bv; best_idx[i] = bi;
        }
        j0 = j1;
    }
    __syncthreads();
    int* out = P.nn + ((size_t)pyi * L.wo + pxi) * L.n1;
    for (int i = tid; i < L.n1; i += NN_THREADS) out[i] = best_idx[i];
}

static size_t search_smem_bytes(const vl3d_loss_desc* L) {
    const int G4 = (3 * L->p + 3) / 4;
    size_t fl = (size_t)4 * 4 * G4 * NN_CF + (size_t)NN_CF * (NN_CF + 1) + (size_t)L->n1 * NN_CF + NN_CF + 2 * (size_t)L->n1;
    return fl * sizeof(float);
}


// ------------------------------------------------------------------------------------------------
// Strip kernel (fast path): one CTA owns a vertical strip of SL patch positions in one patch column and
// sweeps its pixel rows once per candidate chunk.  Vertically overlapping patches (p > s) share their
// row sums: the running sum `cur` of the current group of s rows plus the M = p/s previous group sums
// give every patch as   G_k = R_k + ... + R_{k+M-1} + (first p%s rows of group k+M),
// so each pixel row is reduced ONCE instead of ceil(p/s) times (2.75x fewer FMAs at p=11, s=4).
// Needs all query frames in one register tile (tx_used <= 64) and p/s <= 3; otherwise the
// one-patch-per-CTA kernel above is used.
// ------------------------------------------------------------------------------------------------
struct StripParams {
    vl3d_loss_desc d;
    const float* x;
    const float* y;
    int* nn;
    int groups;     // float4 groups per slab
    int row0, row1; // patch-row range of this launch
    int SL;         // patches per strip
    int nta, ntb;   // thread tile grid: blockDim.x = nta * ntb; x frames XF = 4*nta, chunk frames CF = 4*ntb
};

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 4 : 0;                                // src-size 0 => zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

template <int M, bool VEC>
__global__ void __launch_bounds__(NN_THREADS, 2) patchnn_strip_kernel(const __grid_constant__ StripParams P) {
    extern __shared__ __align__(16) float smem[];
    const vl3d_loss_desc& L = P.d;
    const int G4 = P.groups, NTA = P.nta, NTB = P.ntb, XF = 4 * NTA, CF = 4 * NTB;
    // staging layout: [buffer][frame][G4S] float4, group index fastest.  A thread's operand address is then a
    // per-thread base + an immediate per group (no address arithmetic in the FMA loop), and with G4S odd the
    // lanes of a quarter-warp (consecutive frames, 16*G4S bytes apart) hit disjoint banks.
    const int G4S = G4 | 1;
    const int nthreads = NTA * NTB;
    float4* xs4 = reinterpret_cast<float4*>(smem);                  // [NBUF][XF][G4S]
    float4* ys4 = xs4 + NB * G4S * XF;                         // [NBUF][CF][G4S]
    float* Gs = reinterpret_cast<float*>(ys4 + NB * G4S * CF); // [XF][CF+1]
    float* Ds = Gs + XF * (CF + 1);                                 // [n1][CF+1]
    float* colmin = Ds + (size_t)L.n1 * (CF + 1);                   // [CF]
    float* best_val = colmin + CF;                                  // [SL][n1]
    int* best_idx = reinterpret_cast<int*>(best_val + (size_t)P.SL * L.n1);

    const int tid = threadIdx.x;
    const int pxi = blockIdx.x;
    const int k0 = P.row0 + blockIdx.y * P.SL;
    const int k1 = min(k0 + P.SL, P.row1);
    const int x0 = pxi * L.s;
    const int p = L.p, pt = L.pt, st = L.st, s = L.s;
    const int rem = p - M * s;                                      // rows of the (M+1)-th group used by a patch
    const float inv_d = 1.f / (float)(3 * pt * p * p);
    const int ta = tid / NTB, tb = tid - ta * NTB;
    unsigned xsa[NN_R], ysa[NN_R];                              // shared-space byte addresses of the thread's operands
#pragma unroll
    for (int i = 0; i < NN_R; ++i) {
        xsa[i] = (unsigned)__cvta_generic_to_shared(xs4 + (size_t)(ta + NTA * i) * G4S);
        ysa[i] = (unsigned)__cvta_generic_to_shared(ys4 + (size_t)(tb + NTB * i) * G4S);
    }
    const unsigned xbuf_bytes = (unsigned)(G4S * XF) * 16u, ybuf_bytes = (unsigned)(G4S * CF) * 16u;
    const int tx_used = (L.n1 - 1) * st + pt, ty_used = (L.n2 - 1) * st + pt;
    const int nrows = (k1 - 1 - k0) * s + p;                        // pixel rows swept by this strip
    const int ybase = k0 * s;
    const int cand_per_chunk = (CF - pt) / st + 1;

    for (int i = tid; i < (k1 - k0) * L.n1; i += nthreads) { best_val[i] = INFINITY; best_idx[i] = 0; }

    // asynchronous global -> shared staging of one pixel row of the window (3 channels, all frames):
    // LDGSTS, no register round trip; frames beyond the video are zero-filled.
    //  VEC: 16-byte copies.  Elements are laid out channel-padded (e = c*P4 + dx, P4 = 4*ceil(p/4)), so a
    //       run of p pixels is P4/4 aligned chunks; the tail chunk copies 4*(p%4) bytes and zero-fills.
    //       Needs 16-byte aligned runs (stride % 4 == 0, strides % 4 == 0; host-checked).
    //  else: 4-byte copies, elements packed e = c*p + dx.
    auto stage = [&](int c0, int row, int buf) {
        const int nx = 3 * XF, ny = 3 * CF;
        for (int id = tid; id < nx + ny; id += nthreads) {
            const bool isy = id >= nx;
            const int q = isy ? id - nx : id;
            const int nf = isy ? CF : XF;
            const int c = q / nf, fr = q - c * nf;
            const int gf = isy ? c0 + fr : fr;
            const bool ok = gf < (isy ? ty_used : tx_used);
            const int gfc = ok ? gf : 0;
            const float* src = isy ? P.y + (size_t)gfc * L.y_sf + (size_t)c * L.y_sc + (size_t)(ybase + row) * L.y_sr + x0
                                   : P.x + (size_t)gfc * L.x_sf + (size_t)c * L.x_sc + (size_t)(ybase + row) * L.x_sr + x0;
            if (VEC) {
                const int nch = (p + 3) >> 2;
                float4* d4 = (isy ? ys4 + (size_t)buf * G4S * CF : xs4 + (size_t)buf * G4S * XF) + (size_t)fr * G4S + c * nch;
                for (int j = 0; j < nch; ++j) {
                    const int nval = ok ? min(4, p - 4 * j) * 4 : 0;
                    const unsigned d = (unsigned)__cvta_generic_to_shared(d4 + j);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src + 4 * j), "r"(nval) : "memory");
                }
            } else {
                float* dst = reinterpret_cast<float*>(isy ? ys4 + (size_t)buf * G4S * CF : xs4 + (size_t)buf * G4S * XF);
                float* d0 = dst + (size_t)fr * G4S * 4;
                int e = c * p;
                for (int dx = 0; dx < p; ++dx, ++e) cp_async_f32(d0 + e, src + dx, ok);
            }
        }
        cp_async_commit();
    };
    // packed layout: the padding lanes of the last float4 group never change; zero them once in both buffers
    if (!VEC) {
        const int npad = 4 * G4 - 3 * p;
        for (int id = tid; id < NB * npad * (XF + CF); id += nthreads) {
            const int buf = id / (npad * (XF + CF));
            const int r2 = id - buf * npad * (XF + CF);
            const int e = 3 * p + r2 / (XF + CF);
            const int q = r2 % (XF + CF);
            const bool isy = q >= XF;
            const int fr = isy ? q - XF : q;
            const int nf = isy ? CF : XF;
            float* dst = reinterpret_cast<float*>(isy ? ys4 + (size_t)buf * G4S * CF : xs4 + (size_t)buf * G4S * XF);
            dst[(size_t)fr * G4S * 4 + e] = 0.f;
        }
    }

    for (int j0 = 0; j0 < L.n2;) {
        const int c0 = j0 * st;
        int j1 = j0 + cand_per_chunk;
        if (j1 > L.n2) j1 = L.n2;
        const int cj = j1 - j0;

        float cur[NN_R][NN_R];
        float hist[M > 0 ? M : 1][NN_R][NN_R];                      // hist[0] = most recent complete group
#pragma unroll
        for (int i = 0; i < NN_R; ++i)
#pragma unroll
            for (int j = 0; j < NN_R; ++j) {
                cur[i][j] = 0.f;
#pragma unroll
                for (int m = 0; m < (M > 0 ? M : 1); ++m) hist[m][i][j] = 0.f;
            }

        __syncthreads();                                            // previous chunk's readers are done
        // three staging buffers: rows r+1 and r+2 are in flight while row r is consumed (one row of arithmetic
        // is about as long as a cp.async round trip, so a single row of lookahead left warps waiting here)
        stage(c0, 0, 0);
        if (NB > 2 && nrows > 1) stage(c0, 1, 1);
        int buf = 0;
        for (int row = 0; row < nrows; ++row) {
            if (NB > 2 && row + 1 < nrows) cp_async_wait_1(); else cp_async_wait_all();
            __syncthreads();                                        // row `row` landed; the buffer of row-1 is free
            if (row + NB - 1 < nrows) stage(c0, row + NB - 1, buf >= 1 ? buf - 1 : NB - 1);   // == (buf + NB - 1) % NB
            const int grp = row / s, rin = row - grp * s;           // group of s rows, row inside the group
            if (M > 0 || rin < p) {                                 // (p < s: rows between patches are unused)
                unsigned xa_[NN_R], ya_[NN_R];
#pragma unroll
                for (int i = 0; i < NN_R; ++i) {
                    xa_[i] = xsa[i] + (unsigned)buf * xbuf_bytes;
                    ya_[i] = ysa[i] + (unsigned)buf * ybuf_bytes;
                }
                auto fma_group = [&](int off) {
                    {
                        float4 xa[NN_R], ya[NN_R];
#pragma unroll
                        for (int i = 0; i < NN_R; ++i) xa[i] = lds128(xa_[i] + off);
#pragma unroll
                        for (int j = 0; j < NN_R; ++j) ya[j] = lds128(ya_[j] + off);
#pragma unroll
                        for (int i = 0; i < NN_R; ++i)
#pragma unroll
                            for (int j = 0; j < NN_R; ++j) {
                                float dlt;
                                dlt = xa[i].x - ya[j].x; cur[i][j] = fmaf(dlt, dlt, cur[i][j]);
                                dlt = xa[i].y - ya[j].y; cur[i][j] = fmaf(dlt, dlt, cur[i][j]);
                                dlt = xa[i].z - ya[j].z; cur[i][j] = fmaf(dlt, dlt, cur[i][j]);
                                dlt = xa[i].w - ya[j].w; cur[i][j] = fmaf(dlt, dlt, cur[i][j]);
                            }
                    }
                };
                if (VEC) {                                          // G4 = 3 * ceil(p/4): three groups per trip
                    for (int g = 0; g < G4; g += 3) {
                        fma_group(0); fma_group(16); fma_group(32);
#pragma unroll
                        for (int i = 0; i < NN_R; ++i) { xa_[i] += 48; ya_[i] += 48; }
                    }
                } else {
                    for (int g = 0; g < G4; ++g) {
                        fma_group(0);
#pragma unroll
                        for (int i = 0; i < NN_R; ++i) { xa_[i] += 16; ya_[i] += 16; }
                    }
                }
            }
            // does a patch end on this row?  patch kr (relative) ends at row kr*s + p - 1 = (kr+M)*s + rem - 1
            const bool ends = (rem > 0) ? (rin == rem - 1 && grp >= M) : (rin == s - 1 && grp >= M - 1);
            const int kr = (rem > 0) ? grp - M : grp - (M - 1);
            if (ends && kr < k1 - k0) {
#pragma unroll
                for (int i = 0; i < NN_R; ++i)
#pragma unroll
                    for (int j = 0; j < NN_R; ++j) {
                        float gsum = cur[i][j];
#pragma unroll
                        for (int m = 0; m < M; ++m)
                            if (rem > 0 || m + 1 < M) gsum += hist[m][i][j];   // p == M*s: cur is the M-th group itself
                        Gs[(ta + NTA * i) * (CF + 1) + tb + NTB * j] = gsum;
                    }
                __syncthreads();
                {   // entries id = tid, tid + nthreads, ... <-> (il, jl) = divmod(id, cj), advanced incrementally (a step
                    // of nthreads entries is dq rows + dr columns with one carry) instead of a division per entry
                    const int dq = nthreads / cj, dr = nthreads - dq * cj;
                    int il = tid / cj, jl = tid - il * cj;
                    while (il < L.n1) {
                        const float* gp = Gs + (il * st) * (CF + 1) + jl * st;
                        float sum = 0.f;
                        for (int dt = 0; dt < pt; ++dt) sum += gp[dt * (CF + 2)];
                        Ds[il * (CF + 1) + jl] = sum * inv_d;
                        jl += dr; il += dq;
                        if (jl >= cj) { jl -= cj; ++il; }
                    }
                }
                __syncthreads();
                // The normaliser and the argmin are short reductions over n1 x cj numbers; done by one thread per
                // column / query they are long dependent chains (LDS + IEEE division per step) during which
                // most of the CTA idles, so each is split over PA / PB threads and the partial results are
                // combined in order (same first-minimum / NaN rules as the sequential scan).
                float* part_f = Gs;                                 // Gs is free once Ds is built
                if (L.use_alpha) {
                    const int PA = min(min(8, max(1, nthreads / cj)), XF), RA = (L.n1 + PA - 1) / PA;
                    for (int id = tid; id < cj * PA; id += nthreads) {
                        const int part = id / cj, jl = id - part * cj;
                        const int i1 = min((part + 1) * RA, L.n1);
                        float mn = INFINITY;
                        for (int i = part * RA; i < i1; ++i) {
                            const float vv = Ds[(size_t)i * (CF + 1) + jl];
                            mn = (vv < mn || vv != vv) ? vv : mn;
                        }
                        part_f[part * (CF + 1) + jl] = mn;
                    }
                    __syncthreads();
                    for (int jl = tid; jl < cj; jl += nthreads) {
                        float mn = INFINITY;
                        for (int q = 0; q < PA; ++q) {
                            const float vv = part_f[q * (CF + 1) + jl];
                            mn = (vv < mn || vv != vv) ? vv : mn;
                        }
                        colmin[jl] = L.alpha + mn;
                    }
                    __syncthreads();
                }
                {
                    const int PB = min(8, max(1, nthreads / L.n1)), SB = (cj + PB - 1) / PB;
                    int* part_i = reinterpret_cast<int*>(part_f + PB * L.n1);
                    for (int id = tid; id < L.n1 * PB; id += nthreads) {
                        const int part = id / L.n1, i = id - part * L.n1;
                        const int jb = min((part + 1) * SB, cj);
                        float bv = INFINITY;
                        int bi = -1;
                        for (int jl = part * SB; jl < jb; ++jl) {
                            float vv = Ds[(size_t)i * (CF + 1) + jl];
                            if (L.use_alpha) vv = vv / colmin[jl];
                            const bool better = (vv < bv) || (vv != vv && bv == bv);
                            if (better) { bv = vv; bi = j0 + jl; }
                        }
                        part_f[id] = bv; part_i[id] = bi;
                    }
                    __syncthreads();
                    for (int i = tid; i < L.n1; i += nthreads) {
                        float bv = best_val[kr * L.n1 + i];
                        int bi = best_idx[kr * L.n1 + i];
                        for (int q = 0; q < PB; ++q) {
                            const float vv = part_f[q * L.n1 + i];
                            const bool better = (vv < bv) || (vv != vv && bv == bv);
                            if (better) { bv = vv; bi = part_i[q * L.n1 + i]; }
                        }
                        best_val[kr * L.n1 + i] = bv; best_idx[kr * L.n1 + i] = bi;
                    }
                }
            }
            if (rin == s - 1) {                                     // group complete: shift the history
#pragma unroll
                for (int i = 0; i < NN_R; ++i)
#pragma unroll
                    for (int j = 0; j < NN_R; ++j) {
#pragma unroll
                        for (int m = (M > 0 ? M : 1) - 1; m > 0; --m) hist[m][i][j] = hist[m - 1][i][j];
                        hist[0][i][j] = cur[i][j];
                        cur[i][j] = 0.f;
                    }
            }
            buf = buf + 1 < NB ? buf + 1 : 0;
        }
        j0 = j1;
    }
    __syncthreads();
    for (int id = tid; id < (k1 - k0) * L.n1; id += nthreads) {
        const int kr = id / L.n1, i = id - kr * L.n1;
        P.nn[((size_t)(k0 + kr) * L.wo + pxi) * L.n1 + i] = best_idx[id];
    }
}

static size_t strip_smem_bytes(const vl3d_loss_desc* L, int nta, int ntb, int SL, int nbuf) {
    const int G4 = ((3 * L->p + 3) / 4) | 1, XF = 4 * nta, CF = 4 * ntb;
    size_t fl = (size_t)4 * nbuf * G4 * (XF + CF) + (size_t)XF * (CF + 1) + (size_t)L->n1 * (CF + 1) + CF + 2 * (size_t)SL * L->n1;
    return fl * sizeof(float);
}

}  // namespace vl3d

#include "patchnn_strip8.cuh"
#include "patchnn_diag.cuh"

namespace vl3d {

// ------------------------------------------------------------------------------------------------
// vote / merge + robust loss + its derivative: one thread per pixel of the full x buffer.
//   y2x[c,t',py,px] = mean over covering patches (i,j,k) of y[c, NN_ij[k]*st + t'-k*st, py, px]
// ------------------------------------------------------------------------------------------------
struct VoteParams {
    vl3d_loss_desc d;
    const float* x; const float* xscale; const float* y; const int* nn;
    int rou_kind; float rou, scaling, gcoef;
    int Tx_full, Hfull, Wfull;
    int f0;                     // first frame handled by this launch
    int row0, row1;             // pixel rows whose loss / gradient this call owns (a rank's band; [0, Hfull) normally)
    float n_inv;                // 1 / (elements of the whole fitted crop): partial results add up to the mean
    float* y2x; float* weight; float* grad; double* partials; float* loss;
};

__device__ __forceinline__ void robust(float r, int kind, float rou, float scale, float& val, float& der) {
    if (kind == 1) { val = r * r; der = 2.f * r; return; }
    if (kind == 2) { val = fabsf(r); der = signf(r); return; }
    const float q = r / scale, sq = q * q;
    if (rou == 0.f) { val = log1pf(0.5f * sq); der = (q / scale) / (1.f + 0.5f * sq); return; }
    if (rou == 2.f) { val = 0.5f * sq; der = q / scale; return; }
    const float b = fabsf(rou - 2.f) + 1e-6f;
    const float dd = rou >= 0.f ? rou + 1e-6f : rou - 1e-6f;
    const float base = sq / b + 1.f;                                // >= 1
    // base^(d/2) through MUFU lg2 / ex2 (relative error ~1e-6; two IEEE powf calls per channel were ~40 % of the vote
    // kernel's instructions); the derivative's base^(d/2 - 1) is the same power divided by the base
    const float pw = __powf(base, 0.5f * dd);
    val = (b / dd) * (pw - 1.f) * (scale * 10.f);
    der = 10.f * q * __fdividef(pw, base);
}

constexpr int VOTE_THREADS = 256;

__global__ void __launch_bounds__(VOTE_THREADS) vote_loss_kernel(const __grid_constant__ VoteParams P) {
    const vl3d_loss_desc& L = P.d;
    // frames vary fastest across the grid: the CTAs of one 32x8 pixel tile for all frames are co-resident, so
    // the target-video tile they gather from (all F frames, ~0.8 MB) is read from HBM once and then hit in L2
    const int tf = P.f0 + blockIdx.x;
    const int px = blockIdx.y * 32 + (threadIdx.x & 31);
    const int py = blockIdx.z * 8 + (threadIdx.x >> 5);
    float lsum = 0.f;
    if (px < P.Wfull && py < P.Hfull) {
        const bool inside = px < L.w && py < L.h && tf < L.t && py >= P.row0 && py < P.row1;
        float g[3] = {0.f, 0.f, 0.f};
        if (inside) {
            const int p = L.p, s = L.s, pt = L.pt, st = L.st;
            int iy0 = (py - p + s) / s; if (py - p + 1 <= 0) iy0 = 0;         // ceil((py-p+1)/s) clipped at 0
            int ix0 = (px - p + s) / s; if (px - p + 1 <= 0) ix0 = 0;
            int k0 = (tf - pt + st) / st; if (tf - pt + 1 <= 0) k0 = 0;
            const int iy1 = min(py / s, L.ho - 1), ix1 = min(px / s, L.wo - 1), k1 = min(tf / st, L.n1 - 1);
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            int cnt = 0;
            const float* yb = P.y + (size_t)py * L.y_sr + px;
            for (int iy = iy0; iy <= iy1; ++iy)
                for (int ix = ix0; ix <= ix1; ++ix) {
                    const int* nnp = P.nn + ((size_t)iy * L.wo + ix) * L.n1;
                    for (int k = k0; k <= k1; ++k) {
                        const int fr = __ldg(nnp + k) * st + (tf - k * st);
                        const float* src = yb + (size_t)fr * L.y_sf;
                        v0 += __ldg(src); v1 += __ldg(src + L.y_sc); v2 += __ldg(src + 2 * L.y_sc);
                        ++cnt;
                    }
                }
            const float wgt = fmaxf((float)cnt, 1e-10f);                       // clamp_min(1e-10) (utils_vid.py:228)
            const float m[3] = {v0 / wgt, v1 / wgt, v2 / wgt};
            const float xsc = P.xscale ? __ldg(P.xscale) : 1.f;
            const float n_inv = P.n_inv;
            const size_t fit_plane = (size_t)L.h * L.w;
            const size_t fit_pix = (size_t)py * L.w + px;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xv = __ldg(P.x + (size_t)tf * L.x_sf + (size_t)c * L.x_sc + (size_t)py * L.x_sr + px) * xsc;
                float val, der;
                robust(xv - m[c], P.rou_kind, P.rou, P.scaling, val, der);
                lsum += val;
                g[c] = P.gcoef * xsc * der * n_inv;
                if (P.y2x) P.y2x[((size_t)c * L.t + tf) * fit_plane + fit_pix] = m[c];
            }
            if (P.weight) P.weight[(size_t)tf * fit_plane + fit_pix] = wgt;
        }
        if (P.grad && tf < P.Tx_full) {
            const size_t plane = (size_t)P.Hfull * P.Wfull;
            float* gp = P.grad + (size_t)tf * 3 * plane + (size_t)py * P.Wfull + px;
            gp[0] = g[0]; gp[plane] = g[1]; gp[2 * plane] = g[2];
        }
    }
    __shared__ float s_part[VOTE_THREADS / 32];
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < VOTE_THREADS / 32; ++i) acc += (double)s_part[i];
        P.partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
    }
}

// Same computation with the gathers batched: the loop nest above waits twice per covering patch (NN index, then
// the three target texels it points to: 75 % of its stall samples are on those two scoreboards, issue 36 %).  Here
// the covering patches of one patch row — at most NX columns x NK temporal offsets, compile-time bounds — are
// handled together: all their NN indices are requested first (predicated loads, nothing is fetched for
// combinations that do not exist), then all 3*NX*NK target texels, then the sums.  Up to 9 / 27 loads in flight
// per thread instead of 1 / 3.  The host picks the instantiation from ceil(p/s) and ceil(pt/st).
// IDX32: every element offset into y and nn fits 31 bits (the host checks), so the gather addresses are one IMAD.WIDE
// per load instead of 64-bit multiply / add chains.
template <int NY, int NX, int NK, bool IDX32>
__global__ void __launch_bounds__(VOTE_THREADS) vote_loss_batched_kernel(const __grid_constant__ VoteParams P) {
    const vl3d_loss_desc& L = P.d;
    const int tf = P.f0 + blockIdx.x;
    const int px = blockIdx.y * 32 + (threadIdx.x & 31);
    const int py = blockIdx.z * 8 + (threadIdx.x >> 5);
    float lsum = 0.f;
    if (px < P.Wfull && py < P.Hfull) {
        const bool inside = px < L.w && py < L.h && tf < L.t && py >= P.row0 && py < P.row1;
        float g[3] = {0.f, 0.f, 0.f};
        if (inside) {
            const int p = L.p, s = L.s, pt = L.pt, st = L.st;
            int iy0 = (py - p + s) / s; if (py - p + 1 <= 0) iy0 = 0;
            int ix0 = (px - p + s) / s; if (px - p + 1 <= 0) ix0 = 0;
            int k0 = (tf - pt + st) / st; if (tf - pt + 1 <= 0) k0 = 0;
            const int iy1 = min(py / s, L.ho - 1), ix1 = min(px / s, L.wo - 1), k1 = min(tf / st, L.n1 - 1);
            float v0 = 0.f, v1 = 0.f, v2 = 0.f;
            const int cnt = max(iy1 - iy0 + 1, 0) * max(ix1 - ix0 + 1, 0) * max(k1 - k0 + 1, 0);
            const float* yb = P.y + (size_t)py * L.y_sr + px;
            const size_t nn_row = (size_t)L.wo * L.n1;
            // Patch columns are visited by PHASE (column index mod NX), not by offset from the first covering column:
            // windows of one phase are NX*s >= p pixels apart, so every pixel lies in at most one of them and the
            // 32 adjacent pixels of a warp meet only ~32/(NX*s) + 1 distinct patches per step instead of 32/s — the
            // gathered target texels form ~3 runs of p pixels instead of 8 runs of s pixels (fewer L1 wavefronts per
            // load, and the NN indices are broadcast loads).  The set of patches per pixel is unchanged.
            const int cx = px / s;                                  // last patch column that starts at or before px
#pragma unroll
            for (int jy = 0; jy < NY; ++jy) {
                const bool vy = iy0 + jy <= iy1;
                const int* nny = P.nn + (size_t)(iy0 + jy) * nn_row + k0;
                const int n1i = L.n1, ysf = (int)L.y_sf, ysc = (int)L.y_sc;
                int fr[NX][NK];
#pragma unroll
                for (int jx = 0; jx < NX; ++jx) {
                    const int ix = cx - (cx - jx + NX) % NX;        // largest column <= cx of phase jx
                    const bool vx = vy && ix >= ix0 && ix <= ix1;
#pragma unroll
                    for (int jk = 0; jk < NK; ++jk) {
                        const bool ok = vx && k0 + jk <= k1;
                        int nnv = 0;
                        if (ok) nnv = IDX32 ? __ldg(nny + (ix * n1i + jk)) : __ldg(nny + (size_t)ix * L.n1 + jk);
                        fr[jx][jk] = ok ? nnv * st + (tf - (k0 + jk) * st) : -1;
                    }
                }
                float a0[NX][NK], a1[NX][NK], a2[NX][NK];
#pragma unroll
                for (int jx = 0; jx < NX; ++jx)
#pragma unroll
                    for (int jk = 0; jk < NK; ++jk) {
                        a0[jx][jk] = a1[jx][jk] = a2[jx][jk] = 0.f;
                        if (fr[jx][jk] >= 0) {
                            if (IDX32) {
                                const int o = fr[jx][jk] * ysf;
                                a0[jx][jk] = __ldg(yb + o); a1[jx][jk] = __ldg(yb + (o + ysc)); a2[jx][jk] = __ldg(yb + (o + 2 * ysc));
                            } else {
                                const float* src = yb + (size_t)fr[jx][jk] * L.y_sf;
                                a0[jx][jk] = __ldg(src); a1[jx][jk] = __ldg(src + L.y_sc); a2[jx][jk] = __ldg(src + 2 * L.y_sc);
                            }
                        }
                    }
#pragma unroll
                for (int jx = 0; jx < NX; ++jx)
#pragma unroll
                    for (int jk = 0; jk < NK; ++jk) { v0 += a0[jx][jk]; v1 += a1[jx][jk]; v2 += a2[jx][jk]; }
            }
            const float wgt = fmaxf((float)cnt, 1e-10f);                       // clamp_min(1e-10) (utils_vid.py:228)
            const float m[3] = {v0 / wgt, v1 / wgt, v2 / wgt};
            const float xsc = P.xscale ? __ldg(P.xscale) : 1.f;
            const float n_inv = P.n_inv;
            const size_t fit_plane = (size_t)L.h * L.w;
            const size_t fit_pix = (size_t)py * L.w + px;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xv = __ldg(P.x + (size_t)tf * L.x_sf + (size_t)c * L.x_sc + (size_t)py * L.x_sr + px) * xsc;
                float val, der;
                robust(xv - m[c], P.rou_kind, P.rou, P.scaling, val, der);
                lsum += val;
                g[c] = P.gcoef * xsc * der * n_inv;
                if (P.y2x) P.y2x[((size_t)c * L.t + tf) * fit_plane + fit_pix] = m[c];
            }
            if (P.weight) P.weight[(size_t)tf * fit_plane + fit_pix] = wgt;
        }
        if (P.grad && tf < P.Tx_full) {
            const size_t plane = (size_t)P.Hfull * P.Wfull;
            float* gp = P.grad + (size_t)tf * 3 * plane + (size_t)py * P.Wfull + px;
            gp[0] = g[0]; gp[plane] = g[1]; gp[2 * plane] = g[2];
        }
    }
    __shared__ float s_part[VOTE_THREADS / 32];
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < VOTE_THREADS / 32; ++i) acc += (double)s_part[i];
        P.partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = acc;
    }
}

__global__ void __launch_bounds__(1024) finalize_mean_kernel(const double* partials, int n, double denom, float* out) {
    __shared__ double s[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) out[0] = (float)(acc / denom);
    }
}

// ------------------------------------------------------------------------------------------------
// scale-invariant gain (MPV.py:499-504)
// ------------------------------------------------------------------------------------------------
constexpr int SCALE_BLOCKS = 1184;   // 8 x 148 SMs
constexpr int SCALE_THREADS = 256;

__global__ void __launch_bounds__(SCALE_THREADS) scale_partial_kernel(const float* __restrict__ rgb, int T,
                                                                       const float* __restrict__ res, int F,
                                                                       size_t chw, double* partials) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * SCALE_THREADS + threadIdx.x; i < chw; i += (size_t)gridDim.x * SCALE_THREADS) {
        float sr = 0.f, sx = 0.f;
        for (int f = 0; f < F; ++f) sr += __ldg(res + (size_t)f * chw + i);
        for (int t = 0; t < T; ++t) sx += __ldg(rgb + (size_t)t * chw + i);
        acc += logf((sr / (float)F + 0.01f) / (sx / (float)T + 0.01f));       // torch.log (MPV.py:499-504), not the approximate intrinsic
    }
    __shared__ float s_part[SCALE_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < SCALE_THREADS / 32; ++i) a += (double)s_part[i];
        partials[blockIdx.x] = a;
    }
}

// sharded variant: every rank sums its block of target frames, the (3,H,W) partial sums are
// all-reduced, and the gain is evaluated from the summed target
__global__ void __launch_bounds__(SCALE_THREADS) frame_sum_kernel(const float* __restrict__ v, int n, size_t chw,
                                                                   float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * SCALE_THREADS + threadIdx.x; i < chw; i += (size_t)gridDim.x * SCALE_THREADS) {
        float s = 0.f;
        for (int f = 0; f < n; ++f) s += __ldg(v + (size_t)f * chw + i);
        out[i] = s;
    }
}

__global__ void __launch_bounds__(SCALE_THREADS) scale_partial_presum_kernel(const float* __restrict__ rgb, int T,
                                                                              const float* __restrict__ res_sum, int F,
                                                                              size_t chw, double* partials) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * SCALE_THREADS + threadIdx.x; i < chw; i += (size_t)gridDim.x * SCALE_THREADS) {
        float sx = 0.f;
        for (int t = 0; t < T; ++t) sx += __ldg(rgb + (size_t)t * chw + i);
        acc += logf((__ldg(res_sum + i) / (float)F + 0.01f) / (sx / (float)T + 0.01f));
    }
    __shared__ float s_part[SCALE_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < SCALE_THREADS / 32; ++i) a += (double)s_part[i];
        partials[blockIdx.x] = a;
    }
}

// band-sharded variant: sum of log((mean_F res + .01) / (mean_T rgb + .01)) over the pixel rows [row0, row1) of a
// (.,3,H,W) band buffer (all channels, all columns); the ranks' sums are all-reduced and vl3d_scale_finish turns the
// total into the gain
__global__ void __launch_bounds__(SCALE_THREADS) scale_partial_rows_kernel(const float* __restrict__ rgb, int T,
                                                                            const float* __restrict__ res, int F, int H, int W,
                                                                            int row0, int row1, double* partials) {
    float acc = 0.f;
    const size_t chw = (size_t)3 * H * W, rw = (size_t)(row1 - row0) * W, n = 3 * rw;
    for (size_t j = (size_t)blockIdx.x * SCALE_THREADS + threadIdx.x; j < n; j += (size_t)gridDim.x * SCALE_THREADS) {
        const size_t c = j / rw, r = j - c * rw;
        const size_t i = c * H * W + (size_t)row0 * W + r;
        float sr = 0.f, sx = 0.f;
        for (int f = 0; f < F; ++f) sr += __ldg(res + (size_t)f * chw + i);
        for (int t = 0; t < T; ++t) sx += __ldg(rgb + (size_t)t * chw + i);
        acc += logf((sr / (float)F + 0.01f) / (sx / (float)T + 0.01f));
    }
    __shared__ float s_part[SCALE_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < SCALE_THREADS / 32; ++i) a += (double)s_part[i];
        partials[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(1024) sum_partials_kernel(const double* partials, int n, double* out) {
    __shared__ double s[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) out[0] = acc;
    }
}

__global__ void __launch_bounds__(1024) scale_finalize_kernel(const double* partials, int n, double denom, float* out) {
    __shared__ double s[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) out[0] = ((float)exp(acc / denom) + 3.f) / 4.f;
    }
}

// rgb_pad * scale (MPV.py:504): the search consumes a pre-scaled copy so it can stage with cp.async
__global__ void __launch_bounds__(256) scale_video_kernel(const float* __restrict__ x, const float* __restrict__ xscale,
                                                          float* __restrict__ out, size_t n) {
    const float sc = __ldg(xscale);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = x[i] * sc;
}

// ------------------------------------------------------------------------------------------------
// evaluation metric support (evaluations/NNMSE.py:45-53): mean |Y[NN] - X| per
// (patch position, query): one warp per pair, lanes stride over the 3*pt*p*p patch elements.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_l1_kernel(const vl3d_loss_desc L, const float* __restrict__ x,
                                                       const float* __restrict__ y, const int* __restrict__ nn,
                                                       float* __restrict__ err) {
    const long long pair = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long npairs = (long long)L.ho * L.wo * L.n1;
    if (pair >= npairs) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)(pair % L.n1);
    const long long b = pair / L.n1;
    const int pxi = (int)(b % L.wo), pyi = (int)(b / L.wo);
    const int j = __ldg(nn + pair);
    const int pp = L.p * L.p, d = 3 * L.pt * pp;
    const float* xb = x + (size_t)(i * L.st) * L.x_sf + (size_t)(pyi * L.s) * L.x_sr + pxi * L.s;
    const float* yb = y + (size_t)(j * L.st) * L.y_sf + (size_t)(pyi * L.s) * L.y_sr + pxi * L.s;
    float acc = 0.f;
    for (int e = lane; e < d; e += 32) {
        const int c = e / (L.pt * pp), r = e - c * (L.pt * pp);
        const int dt = r / pp, r2 = r - dt * pp;
        const int dy = r2 / L.p, dx = r2 - dy * L.p;
        const float xv = __ldg(xb + (size_t)dt * L.x_sf + (size_t)c * L.x_sc + (size_t)dy * L.x_sr + dx);
        const float yv = __ldg(yb + (size_t)dt * L.y_sf + (size_t)c * L.y_sc + (size_t)dy * L.y_sr + dx);
        acc += fabsf(yv - xv);
    }
    acc = warp_sum(acc);
    if (lane == 0) err[pair] = acc / (float)d;
}

// uint8 video -> fp32 in [0,1]: `vid / 255` of MVVidPatchDataset (train_3dvid.py:54), done on the device so that only a
// quarter of the bytes cross PCIe.  IEEE division (nvcc default -prec-div=true) = torch's true_divide bit for bit.
// src: `planes` images of H x W bytes with plane / row strides (a crop of a resident video); dst contiguous.
template <bool VEC4>
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int H,
                                                         int W, long long plane_stride, long long row_stride, size_t total) {
    const int Wq = VEC4 ? W / 4 : W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t row = i / Wq;
        const int xq = (int)(i - row * Wq);
        const size_t pl = row / H;
        const int y = (int)(row - pl * H);
        const unsigned char* sp = src + pl * plane_stride + (size_t)y * row_stride;
        if (VEC4) {
            const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(sp) + xq);
            reinterpret_cast<float4*>(dst)[i] = make_float4((float)v.x / 255.f, (float)v.y / 255.f, (float)v.z / 255.f, (float)v.w / 255.f);
        } else {
            dst[i] = (float)__ldg(sp + xq) / 255.f;
        }
    }
}

// to8b (utils.py:17): (255 * clip(x, 0, 1)).astype(uint8), planar (T,3,H,W) float -> (T,H,W,3) uint8
__global__ void __launch_bounds__(256) to8b_kernel(const float* __restrict__ rgb, unsigned char* __restrict__ out, size_t hw,
                                                   size_t total) {
    for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
        const size_t t = idx / hw, pix = idx - t * hw;
        const float* src = rgb + t * 3 * hw + pix;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = fminf(fmaxf(src[c * hw], 0.f), 1.f) * 255.f;
            out[idx * 3 + c] = (unsigned char)v;                 // truncation, as numpy's astype
        }
    }
}

static int validate_desc(const vl3d_loss_desc* L) {
    VL3D_REQUIRE(L != nullptr, VL3D_ENULL, "loss desc is NULL");
    VL3D_REQUIRE(L->p >= 1 && L->pt >= 1 && L->s >= 1 && L->st >= 1, VL3D_EINVAL, "bad patch config");
    VL3D_REQUIRE(L->n1 >= 1 && L->n2 >= 1 && L->ho >= 1 && L->wo >= 1, VL3D_EINVAL, "empty problem");
    VL3D_REQUIRE((L->n1 - 1) * L->st + L->pt <= L->t && (L->n2 - 1) * L->st + L->pt <= L->F, VL3D_EINVAL,
                 "n1/n2 exceed the frame counts");
    VL3D_REQUIRE((L->ho - 1) * L->s + L->p <= L->h && (L->wo - 1) * L->s + L->p <= L->w, VL3D_EINVAL,
                 "patch grid exceeds the crop");
    VL3D_REQUIRE(L->pt <= NN_CF / 2, VL3D_ERANGE, "patcht_size %d too large (max %d)", L->pt, NN_CF / 2);
    VL3D_REQUIRE(L->n1 <= NN_MAX_N1, VL3D_ERANGE, "n1=%d too large (max %d)", L->n1, NN_MAX_N1);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// the two trivial entries of MPMeshVid.losses (MPV.py:135-136): Patch3DMSE / Patch3DAvg (utils_vid.py:437-445)
// on channel-major videos x (3,tx,h,w) and y (3,ty,h,w); value (block partials in double) and dL/dx in one pass.
// ------------------------------------------------------------------------------------------------
constexpr int VMSE_BLOCKS = 1184, VMSE_THREADS = 256;

__global__ void __launch_bounds__(VMSE_THREADS) video_mse_kernel(const float* __restrict__ x, const float* __restrict__ y, int tx,
                                                                 int ty, int frm, size_t hw, float* __restrict__ grad,
                                                                 double* partials) {
    const size_t n = (size_t)3 * frm * hw, per_c = (size_t)frm * hw;
    const float gscale = 2.f / (float)n;
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * VMSE_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * VMSE_THREADS) {
        const size_t c = i / per_c, r = i - c * per_c;             // r = f*hw + pixel
        const size_t xi = c * (size_t)tx * hw + r;
        const float d = __ldg(x + xi) - __ldg(y + c * (size_t)ty * hw + r);
        acc = fmaf(d, d, acc);
        if (grad) grad[xi] = gscale * d;
    }
    __shared__ float s_part[VMSE_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < VMSE_THREADS / 32; ++i) a += (double)s_part[i];
        partials[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(VMSE_THREADS) video_avg_kernel(const float* __restrict__ x, const float* __restrict__ y, int tx,
                                                                 int ty, size_t hw, float* __restrict__ grad, double* partials) {
    const size_t n = (size_t)3 * hw;
    const float gscale = 2.f / ((float)n * (float)tx);
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * VMSE_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * VMSE_THREADS) {
        const size_t c = i / hw, r = i - c * hw;
        const float* xp = x + c * (size_t)tx * hw + r;
        const float* yp = y + c * (size_t)ty * hw + r;
        float sx = 0.f, sy = 0.f;
        for (int f = 0; f < tx; ++f) sx += __ldg(xp + (size_t)f * hw);
        for (int f = 0; f < ty; ++f) sy += __ldg(yp + (size_t)f * hw);
        const float d = sx / (float)tx - sy / (float)ty;
        acc = fmaf(d, d, acc);
        if (grad) {
            float* gp = grad + c * (size_t)tx * hw + r;
            for (int f = 0; f < tx; ++f) gp[(size_t)f * hw] = gscale * d;
        }
    }
    __shared__ float s_part[VMSE_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < VMSE_THREADS / 32; ++i) a += (double)s_part[i];
        partials[blockIdx.x] = a;
    }
}

// tuning knobs (scripts/tune_search.py), read from the environment ONCE per process
struct Knobs { bool tile8, tma, vote_v1, diag; int sl, ntb; };
static const Knobs& knobs() {
    static const Knobs k = [] {
        auto geti = [](const char* n, int d) { const char* e = getenv(n); return e ? atoi(e) : d; };
        return Knobs{geti("VL3D_NN_TILE8", 1) != 0, geti("VL3D_NN_TMA", 1) != 0, geti("VL3D_VOTE_V1", 0) != 0, geti("VL3D_NN_DIAG", 1) != 0,
                     geti("VL3D_NN_SL", 0), geti("VL3D_NN_NTB", 0)};
    }();
    return k;
}

}  // namespace vl3d

using namespace vl3d;

extern "C" int vl3d_patchnn_search(const vl3d_loss_desc* desc, const float* x, const float* y, int32_t row_begin,
                                   int32_t row_end, int32_t* nn_out, void* stream) {
    if (int e = validate_desc(desc)) return e;
    VL3D_REQUIRE(row_begin >= 0 && row_end <= desc->ho && row_begin <= row_end, VL3D_EINVAL, "bad row range [%d,%d)",
                 row_begin, row_end);
    if (row_begin == row_end) return 0;
    VL3D_REQUIRE(x && y && nn_out, VL3D_ENULL, "x / y / nn_out is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const int tx_used = (desc->n1 - 1) * desc->st + desc->pt;
    const int M = desc->p / desc->s;
    const bool strip_ok = M <= 3 && desc->p <= 32;
    if (strip_ok && M >= 1 && desc->p <= 4 && desc->pt == 3 && desc->st == 1 && knobs().diag &&
        ((desc->x_sf | desc->x_sc | desc->x_sr | desc->y_sf | desc->y_sc | desc->y_sr) >= 0)) {
        // small patches (the other-view loss configuration p = 3, s = 2): diagonal sums and arg-min in registers
        // (patchnn_diag.cuh); VL3D_NN_DIAG=0: tuning aid
        const int tail = desc->p;
        void (*kern)(StripParams) = nullptr;
        if (M == 1 && tail == 3) kern = patchnn_diag_kernel<3, 1, 3>;
        else if (M == 1 && tail == 4) kern = patchnn_diag_kernel<3, 1, 4>;
        else if (M == 2 && tail == 4) kern = patchnn_diag_kernel<3, 2, 4>;
        else if (M == 3 && tail == 3) kern = patchnn_diag_kernel<3, 3, 3>;
        StripParams P{};
        P.d = *desc; P.x = x; P.y = y; P.nn = nn_out; P.groups = 3;
        P.row0 = row_begin; P.row1 = row_end;
        P.nta = (desc->n1 + DG_I - 1) / DG_I;
        P.ntb = (desc->n2 + DG_J - 1) / DG_J;
        if (P.ntb > 16) P.ntb = 16;                                 // two CTAs of <= 192 threads per SM (measured: 15.2 vs 16.5 ms with 32)
        if (knobs().ntb >= 1 && knobs().ntb <= 32) P.ntb = min(knobs().ntb, (desc->n2 + DG_J - 1) / DG_J);   // tuning aid
        if (P.nta * P.ntb > DG_MAXT) P.ntb = DG_MAXT / (P.nta > 0 ? P.nta : 1);
        const int rows = row_end - row_begin;
        int SL = knobs().sl >= 2 ? knobs().sl : 24;
        while (SL > 4 && (long long)desc->wo * ((rows + SL - 1) / SL) < 148 * 3) SL >>= 1;   // small grids: more strips
        SL = (rows + (rows + SL - 1) / SL - 1) / ((rows + SL - 1) / SL);
        if (SL > rows) SL = rows;
        P.SL = SL;
        const size_t smem = P.ntb >= 1 ? diag_smem_bytes(desc, P.nta, P.ntb, SL) : (size_t)1 << 30;
        if (kern != nullptr && P.ntb >= 1 && P.nta * P.ntb >= 32 && smem <= 200 * 1024) {
            cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (ce != cudaSuccess) return set_err((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
            dim3 grid(desc->wo, (rows + SL - 1) / SL);
            kern<<<grid, P.nta * P.ntb, smem, st>>>(P);
            return check_launch("patchnn_search(diag)");
        }
    }
    if (strip_ok && tx_used <= 2 * NN_CF) {
        // strip kernels: rows shared between vertically overlapping patches
        {   // 4 x 8 register tiles (patchnn_strip8.cuh) for the common shapes, up to 128 query frames (T = 96 of BASELINE
            // config 5); VL3D_NN_TILE8=0: tuning aid
            const bool tile8 = knobs().tile8;
            const int P4 = (desc->p + 3) / 4 * 4;
            // (a stride that is not a multiple of 4 pixels puts window starts off the 16-byte grid: LDGSTS cannot copy them
            // and the TMA unit raises an illegal-instruction fault for a box that starts at such an element — measured)
            const bool vec = desc->s % 4 == 0 && (3 * desc->p + 3) / 4 == 3 * P4 / 4 &&
                             ((desc->x_sf | desc->x_sc | desc->x_sr | desc->y_sf | desc->y_sc | desc->y_sr) & 3) == 0 &&
                             (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
            if (vec && M >= 1 && M <= 3 && tile8) {
                StripParams P{};
                P.d = *desc; P.x = x; P.y = y; P.nn = nn_out; P.groups = 3 * (P4 / 4);
                P.row0 = row_begin; P.row1 = row_end;
                P.nta = (tx_used + S8_TI - 1) / S8_TI;
                if (P.nta < 2) P.nta = 2;
                const int rows = row_end - row_begin;
                // strip length: long strips share more rows between patches (measured at 720p: 8 / 16 / 24 / 32 patches
                // per strip = 21.3 / 20.2 / 19.8 / 22.0 ms), balanced so that the last strip is not a stub
                int SL = 24;
                if (knobs().sl >= 2) SL = knobs().sl;                   // tuning aid
                if ((long long)desc->wo * ((rows + SL - 1) / SL) < 148 * 6) {
                    // few patch rows (a rank's band of a sharded search): the grid is only a few waves of the 2 x 148
                    // resident CTAs, so pick the strip count that minimises (waves) x (pixel rows a strip sweeps +
                    // its fixed per-strip work, ~8 rows' worth)
                    long long best = 0; int best_sl = 0;
                    for (int strips = 1; strips <= rows; ++strips) {
                        const int sl = (rows + strips - 1) / strips;
                        if (sl < 2 && rows >= 2) break;
                        if (sl > 32) continue;                          // (strip state is sized for <= 32 rows)
                        const long long waves = ((long long)desc->wo * ((rows + sl - 1) / sl) + 295) / 296;
                        const long long cost = waves * (sl * desc->s + desc->p - desc->s + 8);
                        if (best_sl == 0 || cost < best) { best = cost; best_sl = sl; }
                    }
                    SL = best_sl > 0 ? best_sl : rows;
                }
                SL = (rows + (rows + SL - 1) / SL - 1) / ((rows + SL - 1) / SL);
                if (SL > rows) SL = rows;
                // candidate chunk width 8*ntb: fewest sweeps that still leave two CTAs per SM
                int best_ntb = 0; long long best_cost = 0;
                for (int ntb = (desc->pt + S8_TJ - 1) / S8_TJ + 1; ntb <= NN_THREADS / P.nta; ++ntb) {
                    const int cands = (S8_TJ * ntb - desc->pt) / desc->st + 1;
                    const int chunks = (desc->n2 + cands - 1) / cands;
                    const long long cost = (long long)chunks * S8_TJ * ntb * 64 + 4000LL * chunks;
                    if (P.nta * ntb < 64 || strip8_smem_bytes(desc, P.nta, ntb, SL) > 110 * 1024) continue;
                    if (best_ntb == 0 || cost < best_cost) { best_ntb = ntb; best_cost = cost; }
                }
                P.ntb = best_ntb; P.SL = SL;
                const size_t smem = best_ntb ? strip8_smem_bytes(desc, P.nta, P.ntb, SL) : (size_t)1 << 30;
                const int nch = P4 / 4;
                Strip8Params PP;
                // TMA staging (VL3D_NN_TMA=0: tuning aid); the box carries nch | 1 chunks per channel
                const bool fits = best_ntb && smem <= 110 * 1024 && desc->n1 <= S8_TI * P.nta;
                const bool tma = fits && knobs().tma &&
                                 make_video_tmap(&PP.tx, x, desc->x_sf, desc->x_sc, desc->x_sr, tx_used, nch | 1, S8_TI * P.nta) &&
                                 make_video_tmap(&PP.ty, y, desc->y_sf, desc->y_sc, desc->y_sr,
                                                 (desc->n2 - 1) * desc->st + desc->pt, nch | 1, S8_TJ * P.ntb);
                // measured on B200 at 720p (scripts/tune_search.py), 4x8 vs 4x4 tiles.  With TMA staging: p=11 n2=256
                // 17.1 vs 22.0 ms, p=7 14.4 vs 15.8, n2=1024 60.4 vs 78.3, T=24 35.9 vs 35.8 -> always.  With LDGSTS staging
                // the wide chunk only pays when <= 2 sweeps cover the candidates (20.2 vs 22.0; n2=1024: 88 vs 79).
                const int cands8 = best_ntb ? (S8_TJ * best_ntb - desc->pt) / desc->st + 1 : 1;
                const bool few_sweeps = (desc->n2 + cands8 - 1) / cands8 <= 2 && tx_used >= 40;
                if (fits && (tma || few_sweeps)) {
                    const int tail = desc->p - (P4 - 4);
                    PP.P = P;
                    void (*kern)(Strip8Params) = nullptr;
#define VL3D_S8(MM, TT) (tma ? patchnn_strip8_kernel<MM, TT, true> : patchnn_strip8_kernel<MM, TT, false>)
                    if (M == 1) kern = tail == 1 ? VL3D_S8(1, 1) : tail == 2 ? VL3D_S8(1, 2) : tail == 3 ? VL3D_S8(1, 3) : VL3D_S8(1, 4);
                    else if (M == 2) kern = tail == 1 ? VL3D_S8(2, 1) : tail == 2 ? VL3D_S8(2, 2) : tail == 3 ? VL3D_S8(2, 3) : VL3D_S8(2, 4);
                    else kern = tail == 1 ? VL3D_S8(3, 1) : tail == 2 ? VL3D_S8(3, 2) : tail == 3 ? VL3D_S8(3, 3) : VL3D_S8(3, 4);
#undef VL3D_S8
                    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
                    if (ce != cudaSuccess) return set_err((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
                    dim3 grid(desc->wo, (rows + SL - 1) / SL);
                    kern<<<grid, P.nta * P.ntb, smem, st>>>(PP);
                    return check_launch("patchnn_search(strip8)");
                }
            }
        }
    }
    if (strip_ok && tx_used <= NN_CF) {
        StripParams P{};
        P.d = *desc; P.x = x; P.y = y; P.nn = nn_out; P.groups = (3 * desc->p + 3) / 4;
        P.row0 = row_begin; P.row1 = row_end;
        P.nta = (tx_used + 3) / 4;
        if (P.nta < 2) P.nta = 2;
        const int rows = row_end - row_begin;
        int SL = 16;                                                // longer strips share more rows, shorter ones fill the GPU
        while (SL > 2 && (long long)desc->wo * ((rows + SL - 1) / SL) < 148 * 6) SL >>= 1;
        if (SL > rows) SL = rows;
        P.SL = SL;
        // candidate chunk width: 4*ntb frames per sweep; pick the ntb that wastes the fewest frame-sweeps among those
        // whose staging buffers fit in shared memory (few query frames => many threads left for candidates: a wide
        // chunk of a large patch would not fit)
        {
            const int ty_used = (desc->n2 - 1) * desc->st + desc->pt;
            int best_ntb = 0, widest = 0; long long best_cost = 0;
            const int ntb_max = NN_THREADS / P.nta;
            const int ntb_min = (desc->pt + 3) / 4 + 1;
            for (int ntb = ntb_min; ntb <= ntb_max; ++ntb) {
                const int cands = (4 * ntb - desc->pt) / desc->st + 1;
                const int chunks = (desc->n2 + cands - 1) / cands;
                const long long cost = (long long)chunks * 4 * ntb * 64 + 2000LL * chunks;   // + per-chunk overhead
                if (ntb > ntb_min && strip_smem_bytes(desc, P.nta, ntb, SL, NB) > 200 * 1024) break;
                widest = ntb;
                if (P.nta * ntb < 64 && 4 * ntb < ty_used) continue;   // keep at least two warps busy
                if (best_ntb == 0 || cost < best_cost) { best_ntb = ntb; best_cost = cost; }
            }
            if (best_ntb == 0) best_ntb = widest;
            P.ntb = best_ntb;
        }
        const int P4 = (desc->p + 3) / 4 * 4;
        const bool vec = desc->s % 4 == 0 && (3 * desc->p + 3) / 4 == 3 * P4 / 4 &&
                         ((desc->x_sf | desc->x_sc | desc->x_sr | desc->y_sf | desc->y_sc | desc->y_sr) & 3) == 0 &&
                         (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
        const size_t smem = strip_smem_bytes(desc, P.nta, P.ntb, SL, NB);
        VL3D_REQUIRE(smem <= 200 * 1024, VL3D_ERANGE, "patch_size %d / n1 %d need %zu B of shared memory", desc->p,
                     desc->n1, smem);
        void (*kern)(StripParams) =
            vec ? (M == 0 ? patchnn_strip_kernel<0, true> : M == 1 ? patchnn_strip_kernel<1, true>
                 : M == 2 ? patchnn_strip_kernel<2, true> : patchnn_strip_kernel<3, true>)
                : (M == 0 ? patchnn_strip_kernel<0, false> : M == 1 ? patchnn_strip_kernel<1, false>
                 : M == 2 ? patchnn_strip_kernel<2, false> : patchnn_strip_kernel<3, false>);
        cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) return set_err((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
        dim3 grid(desc->wo, (rows + SL - 1) / SL);
        kern<<<grid, P.nta * P.ntb, smem, st>>>(P);
        return check_launch("patchnn_search(strip)");
    }
    const size_t smem = search_smem_bytes(desc);
    VL3D_REQUIRE(smem <= 200 * 1024, VL3D_ERANGE, "patch_size %d / n1 %d need %zu B of shared memory", desc->p, desc->n1, smem);
    cudaError_t ce = cudaFuncSetAttribute(patchnn_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (ce != cudaSuccess) return set_err((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    SearchParams P{};
    P.d = *desc; P.x = x; P.y = y; P.nn = nn_out; P.groups = (3 * desc->p + 3) / 4;
    P.row0 = row_begin;
    dim3 grid(desc->wo, row_end - row_begin);
    patchnn_search_kernel<<<grid, NN_THREADS, smem, st>>>(P);
    return check_launch("patchnn_search");
}

static dim3 vote_grid(const vl3d_loss_desc* L, int Tx_full, int Hfull, int Wfull) {
    return dim3(Tx_full, (Wfull + 31) / 32, (Hfull + 7) / 8);
}

extern "C" int vl3d_vote_partials(int32_t Tx_full, int32_t Hfull, int32_t Wfull) {
    if (Tx_full < 1 || Hfull < 1 || Wfull < 1) return 0;
    return ((Wfull + 31) / 32) * ((Hfull + 7) / 8) * Tx_full;
}

extern "C" int vl3d_vote_loss(const vl3d_loss_desc* desc, const float* x, const float* xscale, const float* y,
                              const int32_t* nn, int32_t rou_kind, float rou, float scaling, float gcoef,
                              int32_t Tx_full, int32_t Hfull, int32_t Wfull, int32_t frame_begin, int32_t frame_end,
                              int32_t row_begin, int32_t row_end, int64_t n_total,
                              float* y2x_out, float* weight_out, float* grad_out, double* partials, float* loss_out,
                              void* stream) {
    if (int e = validate_desc(desc)) return e;
    VL3D_REQUIRE(x && y && nn && partials && loss_out, VL3D_ENULL, "required pointer is NULL");
    VL3D_REQUIRE(Tx_full >= desc->t && Hfull >= desc->h && Wfull >= desc->w, VL3D_EINVAL, "full dims smaller than the crop");
    VL3D_REQUIRE(rou_kind >= 0 && rou_kind <= 2, VL3D_EINVAL, "rou_kind %d", rou_kind);
    VL3D_REQUIRE(frame_begin >= 0 && frame_begin < frame_end && frame_end <= Tx_full, VL3D_EINVAL,
                 "bad frame range [%d,%d)", frame_begin, frame_end);
    VL3D_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= Hfull && n_total >= 0, VL3D_EINVAL,
                 "bad row range [%d,%d) / n_total", row_begin, row_end);
    const double denom = n_total > 0 ? (double)n_total : (double)desc->t * desc->h * desc->w * 3.0;
    dim3 grid = vote_grid(desc, frame_end - frame_begin, Hfull, Wfull);
    const int nblocks = grid.x * grid.y * grid.z;
    VoteParams P{};
    P.d = *desc; P.x = x; P.xscale = xscale; P.y = y; P.nn = nn;
    P.rou_kind = rou_kind; P.rou = rou; P.scaling = scaling; P.gcoef = gcoef;
    P.Tx_full = Tx_full; P.Hfull = Hfull; P.Wfull = Wfull; P.f0 = frame_begin;
    P.row0 = row_begin; P.row1 = row_end; P.n_inv = (float)(1.0 / denom);
    P.y2x = y2x_out; P.weight = weight_out; P.grad = grad_out; P.partials = partials; P.loss = loss_out;
    cudaStream_t st = (cudaStream_t)stream;
    // covering patches per axis: at most ceil(p/s) (space) and ceil(pt/st) (time)
    const int ms = (desc->p + desc->s - 1) / desc->s, mt = (desc->pt + desc->st - 1) / desc->st;
    const bool vote_v1 = knobs().vote_v1;                             // tuning aid: the one-gather-at-a-time kernel
    const bool batched = !vote_v1;
    // 32-bit gather offsets when the target video and the NN map are smaller than 2^31 elements
    const long long y_span = (long long)desc->F * desc->y_sf + 3 * desc->y_sc;
    const bool idx32 = y_span < (1LL << 31) && (long long)desc->wo * desc->n1 + desc->n1 < (1LL << 31);
#define VL3D_VOTE(A, B, C) (idx32 ? vote_loss_batched_kernel<A, B, C, true> : vote_loss_batched_kernel<A, B, C, false>)
    if (batched && ms <= 2 && mt <= 3) VL3D_VOTE(2, 2, 3)<<<grid, VOTE_THREADS, 0, st>>>(P);
    else if (batched && ms <= 3 && mt <= 3) VL3D_VOTE(3, 3, 3)<<<grid, VOTE_THREADS, 0, st>>>(P);
    else if (batched && ms <= 4 && mt <= 3) VL3D_VOTE(4, 4, 3)<<<grid, VOTE_THREADS, 0, st>>>(P);   // p = 15, s = 4
#undef VL3D_VOTE
    else vote_loss_kernel<<<grid, VOTE_THREADS, 0, st>>>(P);
    if (int e = check_launch("vote_loss")) return e;
    finalize_mean_kernel<<<1, 1024, 0, st>>>(partials, nblocks, denom, loss_out);
    return check_launch("vote_finalize");
}

extern "C" int vl3d_video_loss_partials(void) { return VMSE_BLOCKS; }

extern "C" int vl3d_video_loss(int32_t kind, const float* x, const float* y, int32_t tx, int32_t ty, int32_t h, int32_t w,
                               float* grad_out, double* partials, float* loss_out, void* stream) {
    VL3D_REQUIRE(x && y && partials && loss_out, VL3D_ENULL, "video_loss: NULL pointer");
    VL3D_REQUIRE((kind == 0 || kind == 1) && tx >= 1 && ty >= 1 && h >= 1 && w >= 1, VL3D_EINVAL, "video_loss: bad kind / sizes");
    const size_t hw = (size_t)h * w;
    cudaStream_t st = (cudaStream_t)stream;
    if (kind == 0) {
        const int frm = tx < ty ? tx : ty;
        const size_t n = (size_t)3 * frm * hw;
        int blocks = (int)((n + VMSE_THREADS - 1) / VMSE_THREADS);
        if (blocks > VMSE_BLOCKS) blocks = VMSE_BLOCKS;
        if (grad_out && frm < tx) {                                 // frames beyond the common length take no part
            cudaError_t ce = cudaMemsetAsync(grad_out, 0, (size_t)3 * tx * hw * sizeof(float), st);
            if (ce != cudaSuccess) return set_err((int)ce, "video_loss: %s", cudaGetErrorString(ce));
        }
        video_mse_kernel<<<blocks, VMSE_THREADS, 0, st>>>(x, y, tx, ty, frm, hw, grad_out, partials);
        if (int e = check_launch("video_mse")) return e;
        finalize_mean_kernel<<<1, 1024, 0, st>>>(partials, blocks, (double)n, loss_out);
    } else {
        const size_t n = (size_t)3 * hw;
        int blocks = (int)((n + VMSE_THREADS - 1) / VMSE_THREADS);
        if (blocks > VMSE_BLOCKS) blocks = VMSE_BLOCKS;
        video_avg_kernel<<<blocks, VMSE_THREADS, 0, st>>>(x, y, tx, ty, hw, grad_out, partials);
        if (int e = check_launch("video_avg")) return e;
        finalize_mean_kernel<<<1, 1024, 0, st>>>(partials, blocks, (double)n, loss_out);
    }
    return check_launch("video_loss_finalize");
}

extern "C" int vl3d_patch_l1(const vl3d_loss_desc* desc, const float* x, const float* y, const int32_t* nn, float* err_out,
                             void* stream) {
    if (int e = validate_desc(desc)) return e;
    VL3D_REQUIRE(x && y && nn && err_out, VL3D_ENULL, "patch_l1: NULL pointer");
    const long long npairs = (long long)desc->ho * desc->wo * desc->n1;
    patch_l1_kernel<<<(unsigned)((npairs + 7) / 8), 256, 0, (cudaStream_t)stream>>>(*desc, x, y, nn, err_out);
    return check_launch("patch_l1");
}

extern "C" int vl3d_u8_to_unit(const uint8_t* src, float* dst, int32_t planes, int32_t H, int32_t W, int64_t plane_stride,
                               int64_t row_stride, void* stream) {
    VL3D_REQUIRE(src && dst, VL3D_ENULL, "u8_to_unit: NULL pointer");
    VL3D_REQUIRE(planes >= 1 && H >= 1 && W >= 1 && row_stride >= W && plane_stride >= (int64_t)(H - 1) * row_stride + W,
                 VL3D_EINVAL, "u8_to_unit: bad sizes / strides");
    const bool vec = W % 4 == 0 && row_stride % 4 == 0 && plane_stride % 4 == 0 && ((uintptr_t)src & 3) == 0 &&
                     ((uintptr_t)dst & 15) == 0;
    const size_t total = (size_t)planes * H * (vec ? W / 4 : W);
    const size_t blocks = (total + 255) / 256 < (size_t)148 * 32 ? (total + 255) / 256 : (size_t)148 * 32;
    if (vec) u8_to_unit_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, plane_stride, row_stride, total);
    else u8_to_unit_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, plane_stride, row_stride, total);
    return check_launch("u8_to_unit");
}

extern "C" int vl3d_to8b(const float* rgb, uint8_t* out, int32_t T, int32_t H, int32_t W, void* stream) {
    VL3D_REQUIRE(rgb && out, VL3D_ENULL, "to8b: NULL pointer");
    VL3D_REQUIRE(T >= 1 && H >= 1 && W >= 1, VL3D_EINVAL, "to8b: bad sizes");
    const size_t hw = (size_t)H * W, total = hw * T;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    to8b_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rgb, out, hw, total);
    return check_launch("to8b");
}

extern "C" int vl3d_scale_partials(void) { return SCALE_BLOCKS; }

extern "C" int vl3d_scale_video(const float* x, const float* xscale, float* out, int64_t n, void* stream) {
    VL3D_REQUIRE(x && xscale && out, VL3D_ENULL, "scale_video: NULL pointer");
    VL3D_REQUIRE(n >= 0, VL3D_EINVAL, "scale_video: n < 0");
    if (n == 0) return 0;
    size_t blocks = ((size_t)n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_video_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, xscale, out, (size_t)n);
    return check_launch("scale_video");
}

extern "C" int vl3d_frame_sum(const float* v, int32_t n_frames, int64_t chw, float* out, void* stream) {
    VL3D_REQUIRE(v && out, VL3D_ENULL, "frame_sum: NULL pointer");
    VL3D_REQUIRE(n_frames >= 0 && chw >= 1, VL3D_EINVAL, "frame_sum: bad sizes");
    int blocks = (int)(((size_t)chw + SCALE_THREADS - 1) / SCALE_THREADS);
    if (blocks > SCALE_BLOCKS) blocks = SCALE_BLOCKS;
    frame_sum_kernel<<<blocks, SCALE_THREADS, 0, (cudaStream_t)stream>>>(v, n_frames, (size_t)chw, out);
    return check_launch("frame_sum");
}

extern "C" int vl3d_scale_invariant_presum(const float* rgb, int32_t T, const float* res_sum, int32_t F, int32_t H, int32_t W,
                                           double* partials, float* out, void* stream) {
    VL3D_REQUIRE(rgb && res_sum && partials && out, VL3D_ENULL, "required pointer is NULL");
    VL3D_REQUIRE(T >= 1 && F >= 1 && H >= 1 && W >= 1, VL3D_EINVAL, "bad sizes");
    const size_t chw = (size_t)3 * H * W;
    int blocks = (int)((chw + SCALE_THREADS - 1) / SCALE_THREADS);
    if (blocks > SCALE_BLOCKS) blocks = SCALE_BLOCKS;
    cudaStream_t st = (cudaStream_t)stream;
    scale_partial_presum_kernel<<<blocks, SCALE_THREADS, 0, st>>>(rgb, T, res_sum, F, chw, partials);
    if (int e = check_launch("scale_partial_presum")) return e;
    scale_finalize_kernel<<<1, 1024, 0, st>>>(partials, blocks, (double)chw, out);
    return check_launch("scale_finalize");
}

extern "C" int vl3d_scale_log_sum(const float* rgb, int32_t T, const float* res, int32_t F, int32_t H, int32_t W,
                                  int32_t row_begin, int32_t row_end, double* partials, double* sum_out, void* stream) {
    VL3D_REQUIRE(rgb && res && partials && sum_out, VL3D_ENULL, "required pointer is NULL");
    VL3D_REQUIRE(T >= 1 && F >= 1 && H >= 1 && W >= 1 && row_begin >= 0 && row_begin <= row_end && row_end <= H, VL3D_EINVAL,
                 "bad sizes / row range");
    const size_t n = (size_t)3 * (row_end - row_begin) * W;
    int blocks = (int)((n + SCALE_THREADS - 1) / SCALE_THREADS);
    if (blocks > SCALE_BLOCKS) blocks = SCALE_BLOCKS;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    scale_partial_rows_kernel<<<blocks, SCALE_THREADS, 0, st>>>(rgb, T, res, F, H, W, row_begin, row_end, partials);
    if (int e = check_launch("scale_partial_rows")) return e;
    sum_partials_kernel<<<1, 1024, 0, st>>>(partials, blocks, sum_out);
    return check_launch("sum_partials");
}

extern "C" int vl3d_scale_finish(const double* log_sum, int64_t count, float* out, void* stream) {
    VL3D_REQUIRE(log_sum && out, VL3D_ENULL, "required pointer is NULL");
    VL3D_REQUIRE(count >= 1, VL3D_EINVAL, "scale_finish: count < 1");
    scale_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(log_sum, 1, (double)count, out);
    return check_launch("scale_finish");
}

extern "C" int vl3d_scale_invariant(const float* rgb, int32_t T, const float* res, int32_t F, int32_t H, int32_t W,
                                    double* partials, float* out, void* stream) {
    VL3D_REQUIRE(rgb && res && partials && out, VL3D_ENULL, "required pointer is NULL");
    VL3D_REQUIRE(T >= 1 && F >= 1 && H >= 1 && W >= 1, VL3D_EINVAL, "bad sizes");
    const size_t chw = (size_t)3 * H * W;
    int blocks = (int)((chw + SCALE_THREADS - 1) / SCALE_THREADS);
    if (blocks > SCALE_BLOCKS) blocks = SCALE_BLOCKS;
    cudaStream_t st = (cudaStream_t)stream;
    scale_partial_kernel<<<blocks, SCALE_THREADS, 0, st>>>(rgb, T, res, F, chw, partials);
    if (int e = check_launch("scale_partial")) return e;
    scale_finalize_kernel<<<1, 1024, 0, st>>>(partials, blocks, (double)chw, out);
    return check_launch("scale_finalize");
}
