// Strip search, second generation (included by patchnn.cu): 4 x 8 register tiles.
//
// The 4 x 4 strip kernel is bound by the shared-memory pipe, not by arithmetic (ncu: LSU wavefronts 62 %,
// FMA pipe 43 %): a 4 x 4 tile of frame pairs reads 8 operands (LDS.128 = 4 wavefronts each) per 128
// FADD/FFMA, i.e. one wavefront per four arithmetic instructions — exactly the ratio of the SM's one
// wavefront / cycle to its four issue slots / cycle.  Here a thread owns 4 query frames x 8 candidate frames:
// 12 operand loads per 256 arithmetic instructions (-25 % wavefronts), the candidate chunk doubles to 8*NTB
// frames (two sweeps over the query rows for n2 = 256 instead of four: half the query staging and half the
// epilogues), and the padding lane of every channel's last 16-byte chunk is skipped instead of multiplied
// by zero (33 instead of 36 elements per row for p = 11).
//
// Registers: the 32 accumulators of the current row group stay in registers; the history of complete row
// groups (vertical sharing between overlapping patches, touched once per `s` rows) lives in a per-thread
// ring in LOCAL memory (runtime-indexed, L1/L2 backed) instead of 32*M more registers.
// Shared memory: the frame-pair matrix G and the diagonal sums D share one buffer (D is built in registers,
// then written over G), which keeps two CTAs per SM resident with the 136-frame candidate chunk.
#pragma once

namespace vl3d {

constexpr int S8_TI = 4, S8_TJ = 8;

template <int TAIL>
__device__ __forceinline__ void sqdiff_tail(const float4& a, const float4& b, float& acc) {
    float d;
    d = a.x - b.x; acc = fmaf(d, d, acc);
    if (TAIL > 1) { d = a.y - b.y; acc = fmaf(d, d, acc); }
    if (TAIL > 2) { d = a.z - b.z; acc = fmaf(d, d, acc); }
    if (TAIL > 3) { d = a.w - b.w; acc = fmaf(d, d, acc); }
}

struct alignas(64) Strip8Params {
    CUtensorMap tx, ty;                                             // (w, h, c, frame) views of x and y (TMA staging only)
    StripParams P;
};

// M = p / s in {1, 2, 3} (complete row groups per patch), TAIL = p - 4*(ceil(p/4) - 1) in 1..4.
// TMA: a pixel row of the window (3 channels x all frames of the chunk) is fetched by ONE bulk-tensor copy for x and
// one for y — box (4*nchS pixels, 1 row, 3 channels, frames), which lands in exactly the [frame][channel][chunk]
// layout the arithmetic reads (nchS = nch | 1 chunks per channel keep the group stride odd) — instead of ~1700 16-byte LDGSTS per row
// (26 % of the kernel's LSU wavefronts).  The 16-byte tail chunk then holds real pixels beyond the patch instead
// of zeros, which is why the arithmetic skips the padding lane (TAIL) rather than relying on zero fill; frames
// beyond the video and pixels beyond the image are zero-filled by the TMA unit.
template <int M, int TAIL, bool TMA>
__global__ void __launch_bounds__(NN_THREADS, 2) patchnn_strip8_kernel(const __grid_constant__ Strip8Params PP) {
    extern __shared__ __align__(128) float smem8[];                 // (own symbol: TMA destinations need 128-byte alignment)
    float* smem = smem8;
    const StripParams& P = PP.P;
    const vl3d_loss_desc& L = P.d;
    constexpr int NBUF = 2;
    const int G4 = P.groups, NTA = P.nta, NTB = P.ntb, XF = S8_TI * NTA, CF = S8_TJ * NTB;
    const int nch = G4 / 3;                                         // 16-byte chunks per channel run
    const int nchS = nch | 1;                                       // chunks staged per channel (one spare if nch is even)
    const int G4S = 3 * nchS;                                       // odd group stride: conflict-free operand loads
    const int nthreads = NTA * NTB;
    const int xbuf4 = (G4S * XF + 7) & ~7, ybuf4 = (G4S * CF + 7) & ~7;   // float4 per buffer, 128-byte multiples
    float4* xs4 = reinterpret_cast<float4*>(smem);                  // [NBUF][XF][G4S]
    float4* ys4 = xs4 + NBUF * xbuf4;                               // [NBUF][CF][G4S]
    float* Gs = reinterpret_cast<float*>(ys4 + NBUF * ybuf4);       // [XF][CF+1]; D[n1][.] is written over it
    float* part_f = Gs + XF * (CF + 1);                             // [8][CF+1] scratch of the split reductions
    float* colmin = part_f + 8 * (CF + 1);                          // [CF]
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
    unsigned xsa[S8_TI], ysa[S8_TJ];                                // shared-space byte addresses of the operands
#pragma unroll
    for (int i = 0; i < S8_TI; ++i) xsa[i] = (unsigned)__cvta_generic_to_shared(xs4 + (size_t)(ta + NTA * i) * G4S);
#pragma unroll
    for (int j = 0; j < S8_TJ; ++j) ysa[j] = (unsigned)__cvta_generic_to_shared(ys4 + (size_t)(tb + NTB * j) * G4S);
    const unsigned xbuf_bytes = (unsigned)xbuf4 * 16u, ybuf_bytes = (unsigned)ybuf4 * 16u;
    const int tx_used = (L.n1 - 1) * st + pt, ty_used = (L.n2 - 1) * st + pt;
    const int nrows = (k1 - 1 - k0) * s + p;                        // pixel rows swept by this strip
    const int ybase = k0 * s;
    const int cand_per_chunk = (CF - pt) / st + 1;

    for (int i = tid; i < (k1 - k0) * L.n1; i += nthreads) { best_val[i] = INFINITY; best_idx[i] = 0; }

    // LDGSTS staging of one pixel row (3 channels x all frames), 16-byte chunks, channel-padded element order
    // e = c*4*nch + dx; the last chunk of a run copies 4*TAIL bytes and zero-fills (never read beyond a row).
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
            float4* d4 = (isy ? ys4 + (size_t)buf * ybuf4 : xs4 + (size_t)buf * xbuf4) + (size_t)fr * G4S + c * nchS;
            for (int j = 0; j < nch; ++j) {
                const int nval = ok ? min(4, p - 4 * j) * 4 : 0;
                const unsigned d = (unsigned)__cvta_generic_to_shared(d4 + j);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src + 4 * j), "r"(nval) : "memory");
            }
        }
        cp_async_commit();
    };

    __shared__ __align__(8) uint64_t s_full[2];
    if (TMA && tid == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_fence_init();
    }
    unsigned it = 0;                                                // row iterations so far (TMA: buffer = it & 1, phase = it >> 1)
    auto stage_tma = [&](int c0, int row, unsigned slot) {          // thread 0 only
        const int buf = (int)(slot & 1u);
        mbar_arrive_expect_tx(&s_full[buf], (unsigned)(XF + CF) * (unsigned)G4S * 16u);
        tma_load_4d(xs4 + (size_t)buf * xbuf4, &PP.tx, &s_full[buf], x0, ybase + row, 0, 0);
        tma_load_4d(ys4 + (size_t)buf * ybuf4, &PP.ty, &s_full[buf], x0, ybase + row, 0, c0);
    };

    float hist[M][S8_TI * S8_TJ];                                   // ring of complete row groups (local memory)

    for (int j0 = 0; j0 < L.n2;) {
        const int c0 = j0 * st;
        int j1 = j0 + cand_per_chunk;
        if (j1 > L.n2) j1 = L.n2;
        const int cj = j1 - j0;

        float cur[S8_TI][S8_TJ];
#pragma unroll
        for (int i = 0; i < S8_TI; ++i)
#pragma unroll
            for (int j = 0; j < S8_TJ; ++j) cur[i][j] = 0.f;

        __syncthreads();                                            // previous chunk's readers are done (and the barriers exist)
        if (TMA) { if (tid == 0) stage_tma(c0, 0, it); }
        else stage(c0, 0, 0);
        int grp = 0, rin = 0;                                       // group of s rows, row inside the group (no division per row)
        for (int row = 0; row < nrows; ++row, ++it) {
            const int buf = TMA ? (int)(it & 1u) : (row & 1);
            if (TMA) {
                __syncthreads();                                    // everybody is done with row-1: its buffer is free
                if (tid == 0 && row + 1 < nrows) stage_tma(c0, row + 1, it + 1);   // overlaps the arithmetic below
                mbar_wait(&s_full[buf], (it >> 1) & 1u);            // row `row` has landed
            } else {
                cp_async_wait_all();
                __syncthreads();                                    // row `row` landed; buffer buf^1 is free
                if (row + 1 < nrows) stage(c0, row + 1, buf ^ 1);   // overlaps the arithmetic below
            }
            {
                unsigned xa_[S8_TI], ya_[S8_TJ];
#pragma unroll
                for (int i = 0; i < S8_TI; ++i) xa_[i] = xsa[i] + (unsigned)buf * xbuf_bytes;
#pragma unroll
                for (int j = 0; j < S8_TJ; ++j) ya_[j] = ysa[j] + (unsigned)buf * ybuf_bytes;
                for (int c = 0; c < 3; ++c) {
                    for (int j4 = 0; j4 + 1 < nch; ++j4) {          // full chunks of this channel
                        float4 xa[S8_TI], ya[S8_TJ];
#pragma unroll
                        for (int i = 0; i < S8_TI; ++i) xa[i] = lds128(xa_[i]);
#pragma unroll
                        for (int j = 0; j < S8_TJ; ++j) ya[j] = lds128(ya_[j]);
#pragma unroll
                        for (int i = 0; i < S8_TI; ++i)
#pragma unroll
                            for (int j = 0; j < S8_TJ; ++j) sqdiff_tail<4>(xa[i], ya[j], cur[i][j]);
#pragma unroll
                        for (int i = 0; i < S8_TI; ++i) xa_[i] += 16;
#pragma unroll
                        for (int j = 0; j < S8_TJ; ++j) ya_[j] += 16;
                    }
                    {                                               // last chunk: TAIL elements, the rest is padding
                        float4 xa[S8_TI], ya[S8_TJ];
#pragma unroll
                        for (int i = 0; i < S8_TI; ++i) xa[i] = lds128(xa_[i]);
#pragma unroll
                        for (int j = 0; j < S8_TJ; ++j) ya[j] = lds128(ya_[j]);
#pragma unroll
                        for (int i = 0; i < S8_TI; ++i)
#pragma unroll
                            for (int j = 0; j < S8_TJ; ++j) sqdiff_tail<TAIL>(xa[i], ya[j], cur[i][j]);
                        const unsigned adv = 16u * (unsigned)(1 + nchS - nch);      // (skips the spare chunk)
#pragma unroll
                        for (int i = 0; i < S8_TI; ++i) xa_[i] += adv;
#pragma unroll
                        for (int j = 0; j < S8_TJ; ++j) ya_[j] += adv;
                    }
                }
            }
            // does a patch end on this row?  patch kr (relative) ends at row kr*s + p - 1 = (kr+M)*s + rem - 1
            const bool ends = (rem > 0) ? (rin == rem - 1 && grp >= M) : (rin == s - 1 && grp >= M - 1);
            const int kr = (rem > 0) ? grp - M : grp - (M - 1);
            if (ends && kr < k1 - k0) {
                // patch sum = current (partial or M-th) group + the most recent complete groups of the ring
                const int nh = (rem > 0) ? M : M - 1;
#pragma unroll
                for (int i = 0; i < S8_TI; ++i)
#pragma unroll
                    for (int j = 0; j < S8_TJ; ++j) {
                        float gsum = cur[i][j];
                        for (int m = 1; m <= nh; ++m) gsum += hist[(grp - m + 2 * M) % M][i * S8_TJ + j];
                        Gs[(ta + NTA * i) * (CF + 1) + tb + NTB * j] = gsum;
                    }
                __syncthreads();
                // diagonal sums D[il][jl] = sum_dt G[il*st+dt][jl*st+dt] / d, built in registers and written over G
                float* Ds = Gs;
                constexpr int RU = 16;                              // query rows per work unit
                const int NU = (L.n1 + RU - 1) / RU, units = NU * cj;
                if (NU <= 8 && units <= 2 * nthreads) {
                    // Common shapes (n1 <= 128): a work unit is RU rows of ONE candidate column, a thread takes at most two
                    // units.  The column minimum over a unit's rows is formed in registers while D is, so the separate
                    // pass over D for the alpha normaliser (and two CTA barriers) disappear; consecutive threads read
                    // consecutive columns (conflict-free).
                    float dv[2 * RU];
                    int uh[2], uj[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int wu = tid + q * nthreads;
                        const bool on = wu < units;
                        const int h = on ? wu / cj : 0, jl = on ? wu - h * cj : 0;
                        uh[q] = on ? h : -1; uj[q] = jl;
                        float mn = INFINITY;
#pragma unroll
                        for (int r = 0; r < RU; ++r) {
                            const int il = h * RU + r;
                            float sum = 0.f;
                            if (on && il < L.n1) {
                                const float* gp = Gs + (il * st) * (CF + 1) + jl * st;
                                for (int dt = 0; dt < pt; ++dt) sum += gp[dt * (CF + 2)];
                                sum *= inv_d;
                                mn = (sum < mn || sum != sum) ? sum : mn;
                            }
                            dv[q * RU + r] = sum;
                        }
                        if (on && L.use_alpha) part_f[h * (CF + 1) + jl] = mn;
                    }
                    __syncthreads();                                // every G entry has been read; the unit minima are written
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (uh[q] >= 0) {
#pragma unroll
                            for (int r = 0; r < RU; ++r) {
                                const int il = uh[q] * RU + r;
                                if (il < L.n1) Ds[il * (CF + 1) + uj[q]] = dv[q * RU + r];
                            }
                        }
                    }
                    if (L.use_alpha) {
                        for (int jl = tid; jl < cj; jl += nthreads) {
                            float mn = INFINITY;
                            for (int q = 0; q < NU; ++q) {
                                const float vv = part_f[q * (CF + 1) + jl];
                                mn = (vv < mn || vv != vv) ? vv : mn;
                            }
                            colmin[jl] = L.alpha + mn;
                        }
                    }
                    __syncthreads();
                } else {
                    constexpr int DMAX = 32;                            // entries per thread per pass
                    // entry id = base + u*nthreads + tid <-> (il, jl) = divmod(id, cj), advanced incrementally: a step of
                    // nthreads entries is (dq rows, dr columns) with one carry (an integer division per entry cost more
                    // than the three shared-memory reads it addresses)
                    const int dq_step = nthreads / cj, dr_step = nthreads - dq_step * cj;   // divmod(nthreads, cj)
                    int il_b = tid / cj, jl_b = tid - il_b * cj;                             // divmod(tid, cj)
                    for (int base = 0; base < L.n1 * cj; base += nthreads * DMAX) {
                        float dv[DMAX];
                        int il = il_b, jl = jl_b;
    #pragma unroll
                        for (int u = 0; u < DMAX; ++u) {
                            float sum = 0.f;
                            if (il < L.n1) {
                                const float* gp = Gs + (il * st) * (CF + 1) + jl * st;
                                for (int dt = 0; dt < pt; ++dt) sum += gp[dt * (CF + 2)];
                            }
                            dv[u] = sum * inv_d;
                            jl += dr_step; il += dq_step;
                            if (jl >= cj) { jl -= cj; ++il; }
                        }
                        __syncthreads();                                // every G entry of this pass has been read
                        il = il_b; jl = jl_b;
    #pragma unroll
                        for (int u = 0; u < DMAX; ++u) {
                            if (il < L.n1) Gs[il * (CF + 1) + jl] = dv[u];
                            jl += dr_step; il += dq_step;
                            if (jl >= cj) { jl -= cj; ++il; }
                        }
                        il_b = il; jl_b = jl;
                        __syncthreads();
                    }
                    if (L.use_alpha) {
                        const int PA = min(8, max(1, nthreads / cj)), RA = (L.n1 + PA - 1) / PA;
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
            if (rin == s - 1) {                                     // group complete: into the ring, restart cur
                const int slot = grp % M;
#pragma unroll
                for (int i = 0; i < S8_TI; ++i)
#pragma unroll
                    for (int j = 0; j < S8_TJ; ++j) {
                        hist[slot][i * S8_TJ + j] = cur[i][j];
                        cur[i][j] = 0.f;
                    }
                ++grp; rin = 0;
            } else {
                ++rin;
            }
        }
        j0 = j1;
    }
    __syncthreads();
    for (int id = tid; id < (k1 - k0) * L.n1; id += nthreads) {
        const int kr = id / L.n1, i = id - kr * L.n1;
        P.nn[((size_t)(k0 + kr) * L.wo + pxi) * L.n1 + i] = best_idx[id];
    }
}

static size_t strip8_smem_bytes(const vl3d_loss_desc* L, int nta, int ntb, int SL) {
    const int G4S = 3 * (((L->p + 3) / 4) | 1), XF = S8_TI * nta, CF = S8_TJ * ntb;
    size_t fl = (size_t)4 * 2 * (((G4S * XF + 7) & ~7) + ((G4S * CF + 7) & ~7)) + (size_t)XF * (CF + 1) + (size_t)8 * (CF + 1) + CF + 2 * (size_t)SL * L->n1;
    return fl * sizeof(float);
}

// (w, h, c, frame) tensor map of a planar video with element strides (frame, channel, row); box = (4*nch, 1, 3, nframes)
static bool make_video_tmap(CUtensorMap* out, const float* base, long long sf, long long sc, long long sr, int frames, int nch,
                            int box_frames) {
    vl3d_encode_tiled_fn enc = tma_encoder();
    if (enc == nullptr || box_frames > 256 || frames < 1 || sr < 4 * nch || sc < sr || sf < 3 * sc) return false;
    const cuuint64_t gdim[4] = {(cuuint64_t)sr, (cuuint64_t)(sc / sr), 3, (cuuint64_t)frames};
    const cuuint64_t gstr[3] = {(cuuint64_t)sr * 4, (cuuint64_t)sc * 4, (cuuint64_t)sf * 4};
    const cuuint32_t box[4] = {(cuuint32_t)(4 * nch), 1, 3, (cuuint32_t)box_frames};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (gstr[2] >= ((cuuint64_t)1 << 40)) return false;
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vl3d
