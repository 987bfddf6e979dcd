// Strip search for SMALL patches (p <= 4: one 16-byte chunk per channel and pixel row), included by patchnn.cu.
//
// The other-view loss configuration of the reference (configs/mpv_base.txt:60-68: patch_size 3, stride 2, patcht 3,
// alpha = None) is what 8 of 9 training views run.  Its patches hold 27 values per frame pair, so in the 4 x 4 / 4 x 8
// strip kernels the per-position epilogue — frame-pair matrix G to shared memory, diagonal sums D[i][j] = sum_dt
// G[i+dt][j+dt] read back from it, arg-min scans over D in shared memory, ~10 shared-memory operations per entry —
// costs more than the arithmetic (28 ms at 720p for 1.6e11 lane-operations, 5.6x less efficient than the p = 11 search).
//
// Here a thread owns a CONTIGUOUS block of DG_I query positions x DG_J candidates and accumulates the
// (DG_I + pt - 1) x (DG_J + pt - 1) frame pairs that its diagonals need (6 x 10 for pt = 3: 1.9x the pairs, still only
// 67 lane-operations per D entry), so D is formed in registers; the column minimum (alpha normaliser,
// utils_vid.py:118,133-134) and the arg-min (utils_vid.py:122-142, first minimum) reduce over a thread's registers first
// and only DG_I + DG_J partial results per thread cross shared memory.  Rows are shared between vertically overlapping
// patches exactly as in the other strip kernels (ring of complete row groups in local memory).  Window starts may sit
// on any 4-byte boundary (stride 2): rows are staged with 4-byte LDGSTS.
#pragma once

namespace vl3d {

constexpr int DG_I = 4, DG_J = 8;
constexpr int DG_MAXT = 384;                                        // threads per CTA (168 registers per thread)
constexpr int DG_RB = 1;                                            // pixel rows staged (and swept) per CTA barrier (2: measured slower, 17.1 vs 15.2 ms)

template <int PT, int M, int TAIL>
__global__ void __launch_bounds__(DG_MAXT, 1) patchnn_diag_kernel(const __grid_constant__ StripParams P) {
    extern __shared__ __align__(16) float smemd[];
    const vl3d_loss_desc& L = P.d;
    constexpr int GA = DG_I + PT - 1, GB = DG_J + PT - 1;
    const int NTA = P.nta, NTB = P.ntb, nthreads = NTA * NTB;
    const int XFR = DG_I * NTA + PT - 1;                            // x frames staged per row
    const int CJ = DG_J * NTB;                                      // candidates per sweep
    const int YFR = CJ + PT - 1;                                    // y frames staged per row and sweep
    const int YIDX = YFR + (YFR >> 3) + 1;                          // one spare slot per 8 frames: lanes 9 slots apart => no bank conflicts
    float4* xs4 = reinterpret_cast<float4*>(smemd);                 // [2][DG_RB][3][XFR]
    float4* ys4 = xs4 + 2 * DG_RB * 3 * XFR;                        // [2][DG_RB][3][YIDX]
    float* part = reinterpret_cast<float*>(ys4 + 2 * DG_RB * 3 * YIDX);   // [NTA][CJ] column-minimum partials
    float* colmin = part + NTA * CJ;                                // [CJ]
    float* pbv = colmin + CJ;                                       // [DG_I*NTA][NTB] arg-min partials: value
    int* pbi = reinterpret_cast<int*>(pbv + DG_I * NTA * NTB);      //                                     index
    float* best_val = reinterpret_cast<float*>(pbi + DG_I * NTA * NTB);   // [SL][n1]
    int* best_idx = reinterpret_cast<int*>(best_val + (size_t)P.SL * L.n1);

    const int tid = threadIdx.x;
    const int ta = tid / NTB, tb = tid - ta * NTB;
    const int pxi = blockIdx.x;
    const int k0 = P.row0 + blockIdx.y * P.SL;
    const int k1 = min(k0 + P.SL, P.row1);
    const int x0 = pxi * L.s;
    const int p = L.p, s = L.s;
    const int rem = p - M * s;
    const float inv_d = 1.f / (float)(3 * PT * p * p);
    const int tx_used = L.n1 - 1 + PT, ty_used = L.n2 - 1 + PT;    // (temporal stride 1)
    const int nrows = (k1 - 1 - k0) * s + p;
    const int ybase = k0 * s;

    for (int i = tid; i < (k1 - k0) * L.n1; i += nthreads) { best_val[i] = INFINITY; best_idx[i] = 0; }

    // one pixel row: p floats per (frame, channel) slot; frames beyond the videos are zero-filled
    // (a thread copies all three channels of a frame: no index divisions; 8-byte copies where the window start allows)
    const bool pair_ok = ((x0 | (int)L.x_sr | (int)L.x_sc | (int)L.x_sf | (int)L.y_sr | (int)L.y_sc | (int)L.y_sf) & 1) == 0 &&
                         (((uintptr_t)P.x | (uintptr_t)P.y) & 7) == 0;
    auto stage = [&](int c0, int row0, int buf) {
        const int nr = min(DG_RB, nrows - row0);
        for (int idr = tid; idr < (XFR + YFR) * nr; idr += nthreads) {
            const int rr = DG_RB == 1 ? 0 : idr / (XFR + YFR), id = idr - rr * (XFR + YFR);
            const int row = row0 + rr;
            const bool isy = id >= XFR;
            const int fr = isy ? id - XFR : id;
            const int gf = isy ? c0 + fr : fr;
            const bool ok = gf < (isy ? ty_used : tx_used);
            const int gfc = ok ? gf : 0;
            const float* src = isy ? P.y + (size_t)gfc * L.y_sf + (size_t)(ybase + row) * L.y_sr + x0
                                   : P.x + (size_t)gfc * L.x_sf + (size_t)(ybase + row) * L.x_sr + x0;
            const size_t sc = isy ? (size_t)L.y_sc : (size_t)L.x_sc;
            float4* d4 = isy ? ys4 + (size_t)((buf * DG_RB + rr) * 3) * YIDX + fr + (fr >> 3)
                             : xs4 + (size_t)((buf * DG_RB + rr) * 3) * XFR + fr;
            const int dstep = isy ? YIDX : XFR;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float* d = reinterpret_cast<float*>(d4 + c * dstep);
                const float* sp = src + c * sc;
                if (pair_ok) {
                    const unsigned da = (unsigned)__cvta_generic_to_shared(d);
                    const int nb = ok ? 8 : 0;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da), "l"(sp), "r"(nb) : "memory");
                    if (TAIL == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da + 8u), "l"(sp + 2), "r"(nb) : "memory");
                    else if (TAIL == 3) cp_async_f32(d + 2, sp + 2, ok);
                } else {
#pragma unroll
                    for (int e = 0; e < TAIL; ++e) cp_async_f32(d + e, sp + e, ok);
                }
            }
        }
        cp_async_commit();
    };

    const unsigned xs_base = (unsigned)__cvta_generic_to_shared(xs4 + DG_I * ta);
    const unsigned ys_base = (unsigned)__cvta_generic_to_shared(ys4 + DG_J * tb + ((DG_J * tb) >> 3));
    // slot of y frame DG_J*tb + b relative to ys_base: b + carries into the next blocks of 8 (DG_J = 8: one per block)
    float hist[M][GA * GB];                                         // ring of complete row groups (local memory)

    for (int j0 = 0; j0 < L.n2; j0 += CJ) {
        const int cj = min(CJ, L.n2 - j0);
        float cur[GA][GB];
#pragma unroll
        for (int a = 0; a < GA; ++a)
#pragma unroll
            for (int b = 0; b < GB; ++b) cur[a][b] = 0.f;
        __syncthreads();                                            // previous sweep's readers are done
        stage(j0, 0, 0);
        int grp = 0, rin = 0;                                       // group of s rows, row inside the group
        for (int row0 = 0; row0 < nrows; row0 += DG_RB) {
          const int buf = (row0 / DG_RB) & 1;
          cp_async_wait_all();
          __syncthreads();                                          // rows row0.. landed; buffer buf^1 is free
          if (row0 + DG_RB < nrows) stage(j0, row0 + DG_RB, buf ^ 1);   // overlaps the arithmetic below
          for (int rr = 0; rr < DG_RB && row0 + rr < nrows; ++rr) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float4 xa[GA];
                const unsigned xb = xs_base + (unsigned)(((buf * DG_RB + rr) * 3 + c) * XFR) * 16u;
                const unsigned yb = ys_base + (unsigned)(((buf * DG_RB + rr) * 3 + c) * YIDX) * 16u;
#pragma unroll
                for (int a = 0; a < GA; ++a) xa[a] = lds128(xb + 16u * a);
#pragma unroll
                for (int b = 0; b < GB; ++b) {
                    const float4 ya = lds128(yb + 16u * (b + (b >> 3)));
#pragma unroll
                    for (int a = 0; a < GA; ++a) sqdiff_tail<TAIL>(xa[a], ya, cur[a][b]);
                }
            }
            const bool ends = (rem > 0) ? (rin == rem - 1 && grp >= M) : (rin == s - 1 && grp >= M - 1);
            const int kr = (rem > 0) ? grp - M : grp - (M - 1);
            if (ends && kr < k1 - k0) {
                const int nh = (rem > 0) ? M : M - 1;
                // D[i][j] = sum_dt G[i+dt][j+dt] / d with G = current group + the ring (utils_vid.py:72-100 after unfold)
                float D[DG_I][DG_J];
#pragma unroll
                for (int i = 0; i < DG_I; ++i)
#pragma unroll
                    for (int j = 0; j < DG_J; ++j) {
                        float sum = 0.f;
#pragma unroll
                        for (int dt = 0; dt < PT; ++dt) {
                            float g = cur[i + dt][j + dt];
                            for (int m = 1; m <= nh; ++m) g += hist[(grp - m + 2 * M) % M][(i + dt) * GB + j + dt];
                            sum += g;
                        }
                        const bool valid = DG_I * ta + i < L.n1 && DG_J * tb + j < cj;
                        D[i][j] = valid ? sum * inv_d : INFINITY;
                    }
                if (L.use_alpha) {
                    // column minima over ALL query positions: registers, then NTA partials per candidate
#pragma unroll
                    for (int j = 0; j < DG_J; ++j) {
                        float mn = INFINITY;
#pragma unroll
                        for (int i = 0; i < DG_I; ++i) {
                            const float vv = D[i][j];
                            if (DG_I * ta + i < L.n1) mn = (vv < mn || vv != vv) ? vv : mn;
                        }
                        part[ta * CJ + DG_J * tb + j] = mn;
                    }
                    __syncthreads();
                    for (int jl = tid; jl < cj; jl += nthreads) {
                        float mn = INFINITY;
                        for (int q = 0; q < NTA; ++q) {
                            const float vv = part[q * CJ + jl];
                            mn = (vv < mn || vv != vv) ? vv : mn;
                        }
                        colmin[jl] = L.alpha + mn;
                    }
                    __syncthreads();
                }
                // per-thread first minimum with torch.argmin's NaN rule (a NaN beats everything, the first one wins): the
                // values are >= 0 or NaN, so as signed integers their bit patterns order like the values, and a NaN is
                // mapped to -1.  Entries beyond the candidates are +inf (above) and never beat a real one.
#pragma unroll
                for (int i = 0; i < DG_I; ++i) {
                    int bk = 0x7f800000, bj = -1;
#pragma unroll
                    for (int j = 0; j < DG_J; ++j) {
                        float vv = D[i][j];
                        if (L.use_alpha && DG_J * tb + j < cj) vv = vv / colmin[DG_J * tb + j];
                        const int key = (vv != vv) ? -1 : __float_as_int(vv);
                        if (key < bk) { bk = key; bj = j; }
                    }
                    pbv[(DG_I * ta + i) * NTB + tb] = bk == -1 ? __int_as_float(0x7fc00000) : __int_as_float(bk);
                    pbi[(DG_I * ta + i) * NTB + tb] = bj >= 0 ? j0 + DG_J * tb + bj : -1;
                }
                __syncthreads();
                for (int i = tid; i < L.n1; i += nthreads) {
                    float bv = best_val[kr * L.n1 + i];
                    int bi = best_idx[kr * L.n1 + i];
                    for (int q = 0; q < NTB; ++q) {                 // ascending candidates: the first minimum wins
                        const float vv = pbv[i * NTB + q];
                        const bool better = (vv < bv) || (vv != vv && bv == bv);
                        if (better && pbi[i * NTB + q] >= 0) { bv = vv; bi = pbi[i * NTB + q]; }
                    }
                    best_val[kr * L.n1 + i] = bv; best_idx[kr * L.n1 + i] = bi;
                }
                // (the partial arrays are rewritten at the next patch end, at least one row barrier from here)
            }
            if (rin == s - 1) {                                     // group complete: into the ring, restart cur
                const int slot = grp % M;
#pragma unroll
                for (int a = 0; a < GA; ++a)
#pragma unroll
                    for (int b = 0; b < GB; ++b) {
                        hist[slot][a * GB + b] = cur[a][b];
                        cur[a][b] = 0.f;
                    }
                ++grp; rin = 0;
            } else {
                ++rin;
            }
          }
        }
    }
    __syncthreads();
    for (int id = tid; id < (k1 - k0) * L.n1; id += nthreads) {
        const int kr = id / L.n1, i = id - kr * L.n1;
        P.nn[((size_t)(k0 + kr) * L.wo + pxi) * L.n1 + i] = best_idx[id];
    }
}

static size_t diag_smem_bytes(const vl3d_loss_desc* L, int nta, int ntb, int SL) {
    const int XFR = DG_I * nta + L->pt - 1, CJ = DG_J * ntb, YFR = CJ + L->pt - 1, YIDX = YFR + (YFR >> 3) + 1;
    size_t fl = (size_t)4 * 2 * DG_RB * 3 * (XFR + YIDX) + (size_t)nta * CJ + CJ + 2 * (size_t)DG_I * nta * ntb + 2 * (size_t)SL * L->n1;
    return fl * sizeof(float);
}

}  // namespace vl3d
