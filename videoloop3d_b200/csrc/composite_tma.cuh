// TMA-staged render (included by composite.cu after composite_lean.cuh).
//
// The per-thread render is capped by its own load structure: every pixel issues four LDG.128 per (plane,
// frame), i.e. 16 KB of L1 requests for the 4.75 KB of unique texels a 32x8 tile touches, and a probe that
// only sums the four taps tops out at ~4.5 TB/s (profiles/README.md).  When every plane is one contiguous
// rectangle of the dynamic atlas (VL3D_VIEW_RECT_PLANES: the dense layout of MPV.py:75-81) the footprint of a
// screen tile on a plane is a rectangle too, so ONE elected thread fetches it with `cp.async.bulk.tensor`
// (TMA) — a 40x12-texel box per (tile, plane, frame), unique bytes only, no registers, no L1 tag traffic —
// into a three-stage shared-memory ring guarded by full/empty mbarriers, and the eight consumer warps
// (thread = pixel, exactly the arithmetic of composite_render_kernel) filter from shared memory.
// A tap that falls outside the box (footprint larger than 40x12: strong minification, or a rounding
// difference at the box edge) is read from global memory instead, so results never depend on the box:
// the output is bit-identical to composite_render_kernel.
#pragma once

namespace vl3d {

template <int TF, int TMA_STAGES>
__global__ void __launch_bounds__(TMA_THREADS, (TMA_STAGES * TF * TMA_BOX_BYTES <= 70 * 1024) ? 3 : 2) composite_render_tma_kernel(const __grid_constant__ TmaRenderParams P) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    float4* tiles = reinterpret_cast<float4*>(tma_smem);                                   // [STAGES][TF][BH][BW]
    uint64_t* full = reinterpret_cast<uint64_t*>(tma_smem + TMA_STAGES * TF * TMA_BOX_BYTES);
    uint64_t* empty = full + TMA_STAGES;
    int4* boxinfo = reinterpret_cast<int4*>(empty + TMA_STAGES);                           // [D] (bx0, by0, valid, -)
    const CompositeParams& p = P.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (BX * BY) / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    const int H = p.view.H, W = p.view.W, D = p.view.D;
    const int t0 = p.tb + blockIdx.z * TF;
    const int tile_x0 = blockIdx.x * BX, tile_y0 = blockIdx.y * BY;
    const float qwf = (float)p.view.qw, qhf = (float)p.view.qh;

    if (warp == (BX * BY) / 32) {
        // ===== producer warp: lane d computes the tile's texel box on plane d (all planes at once: the boxes
        // depend on nothing), then lane 0 streams the TMA loads through the ring =====
        if (lane < D) {
            const int d = lane;
            const float cu[2] = {(float)tile_x0 + 0.5f - p.view.cx, (float)min(tile_x0 + BX - 1, W - 1) + 0.5f - p.view.cx};
            const float cv[2] = {(float)tile_y0 + 0.5f - p.view.cy, (float)min(tile_y0 + BY - 1, H - 1) + 0.5f - p.view.cy};
            const float* h = &p.view.hom[d * 9];
            float lxmin = 3e38f, lymin = 3e38f, gxmin = 3e38f, gxmax = -3e38f, gymin = 3e38f, gymax = -3e38f;
            bool front = true;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float u = cu[c & 1], v = cv[c >> 1];
                const float w = fmaf(h[6], u, fmaf(h[7], v, h[8]));
                front = front && (w > 0.f);
                float inv = rcp_ftz(w);
                inv = inv * fmaf(-w, inv, 2.f);
                const float gx = fmaf(h[0], u, fmaf(h[1], v, h[2])) * inv, gy = fmaf(h[3], u, fmaf(h[4], v, h[5])) * inv;
                gxmin = fminf(gxmin, gx); gxmax = fmaxf(gxmax, gx); gymin = fminf(gymin, gy); gymax = fmaxf(gymax, gy);
                const float gxc = fminf(fmaxf(gx, 0.f), qwf), gyc = fminf(fmaxf(gy, 0.f), qhf);
                const int qx = min((int)gxc, p.view.qw - 1), qy = min((int)gyc, p.view.qh - 1);
                const float4* qp = reinterpret_cast<const float4*>(&p.quads[(d * p.view.qh + qy) * p.view.qw + qx]);
                const float4 qa = __ldg(qp);
                const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
                lxmin = fminf(lxmin, (float)qb.x + fmaf(gxc - (float)qx, qa.z, qa.x));
                lymin = fminf(lymin, (float)qb.y + fmaf(gyc - (float)qy, qa.w, qa.y));
            }
            const bool valid = front && gxmax > 0.f && gxmin < qwf && gymax > 0.f && gymin < qhf;
            boxinfo[d] = make_int4(valid ? (int)floorf(lxmin) : 0, valid ? (int)floorf(lymin) : 0, valid ? 1 : 0, 0);
        }
        __syncwarp();
        if (lane != 0) return;
        for (int d = 0; d < D; ++d) {
            const int s = d % TMA_STAGES;
            const unsigned ph = (unsigned)(d / TMA_STAGES) & 1u;
            mbar_wait(&empty[s], ph ^ 1u);                          // the consumers have released this stage
            const int4 bi = boxinfo[d];
            if (bi.z != 0) {
                mbar_arrive_expect_tx(&full[s], TF * TMA_BOX_BYTES);
#pragma unroll
                for (int f = 0; f < TF; ++f)
                    tma_load_3d(tiles + (size_t)(s * TF + f) * (TMA_BOX_BYTES / 16), &P.tmap, &full[s], bi.x * 4, bi.y, t0 + f);
            } else {
                mbar_arrive(&full[s]);
            }
        }
        return;
    }

    // ===== consumers: thread = pixel =====
    const int px = tile_x0 + threadIdx.x % BX, py = tile_y0 + threadIdx.x / BX;
    const bool active = px < W && py < H;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;
    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* fb[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) fb[f] = opaque_ptr(p.atlas_dyn + (size_t)(t0 + f) * dyn_frame);
    const float4* sb = opaque_ptr(p.atlas_sta);
    float Tr[TF], cr[TF], cg[TF], cb[TF], ca[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) { Tr[f] = 1.f; cr[f] = cg[f] = cb[f] = ca[f] = 0.f; }
    int nhit = 0;
    for (int d = 0; d < D; ++d) {
        const int s = d % TMA_STAGES;
        const unsigned ph = (unsigned)(d / TMA_STAGES) & 1u;
        mbar_wait(&full[s], ph);                                    // the plane's boxes have landed
        float gx, gy;
        if (active && plane_grid_lean(&p.view.hom[d * 9], u, v, qwf, qhf, gx, gy)) {
            int qx, qy;
            const float4* qp = quad_at(p, d, gx, gy, qx, qy);
            const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
            if (qb.z != 0) {
                const GeoXY t = geoxy_from_quad(p, qp, qb, qx, qy, gx, gy);
                ++nhit;
                float4 val[TF];
                if (t.g.kind == 2) {
                    const int4 bi = boxinfo[d];
                    const int lx0 = t.cx0 - bi.x, lx1 = t.cx1 - bi.x, ly0 = t.cy0 - bi.y, ly1 = t.cy1 - bi.y;
                    if (bi.z != 0 && lx0 >= 0 && lx1 < TMA_BW && ly0 >= 0 && ly1 < TMA_BH) {
                        const float4* tb = tiles + (size_t)(s * TF) * (TMA_BOX_BYTES / 16);
                        const int a00 = ly0 * TMA_BW + lx0, a10 = ly0 * TMA_BW + lx1, a01 = ly1 * TMA_BW + lx0, a11 = ly1 * TMA_BW + lx1;
#pragma unroll
                        for (int f = 0; f < TF; ++f) {
                            const float4* tf = tb + f * (TMA_BOX_BYTES / 16);
                            val[f] = filter_taps(tf[a00], tf[a10], tf[a01], tf[a11], t.g);
                        }
                    } else {
#pragma unroll
                        for (int f = 0; f < TF; ++f) val[f] = sample_lean(fb[f], t.g);
                    }
                } else {
                    const float4 sv = sample_lean(sb, t.g);               // static tile: same for all frames (MPV.py:445)
#pragma unroll
                    for (int f = 0; f < TF; ++f) val[f] = sv;
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
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);                      // this warp is done with the stage
    }
    if (!active) return;
    if (p.hits_out != nullptr && t0 == 0) p.hits_out[py * W + px] = nhit;
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = t0 + f;
        float* o = p.rgb_out + (size_t)t * 3 * plane + pix;
        o[0] = cr[f]; o[plane] = cg[f]; o[2 * plane] = cb[f];
        if (t < p.pad) {                                          // loop pad: cat(rgb, rgb[:pt-1]) (MPV.py:490-492)
            float* o2 = p.rgb_out + (size_t)(p.T + t) * 3 * plane + pix;
            o2[0] = cr[f]; o2[plane] = cg[f]; o2[2 * plane] = cb[f];
        }
        if (p.alpha_out) p.alpha_out[(size_t)t * plane + pix] = ca[f];
    }
}

// Launches the TMA render for frames [0, TF * (T / TF)); returns false (nothing launched) if TMA is unavailable.
template <int TF, int TMA_STAGES>
static bool launch_render_tma(const CompositeParams& p, const float* atlas_dyn, int T, cudaStream_t st) {
    const int nz = T / TF;
    if (nz == 0) return false;
    TmaRenderParams P;
    if (!make_atlas_tmap(&P.tmap, p.view, atlas_dyn, T)) return false;
    P.p = p;
    P.p.tb = 0;
    const size_t smem = (size_t)TMA_STAGES * TF * TMA_BOX_BYTES + 2 * TMA_STAGES * sizeof(uint64_t) + VL3D_MAX_PLANES * sizeof(int4) + 64;
    if (cudaFuncSetAttribute(composite_render_tma_kernel<TF, TMA_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    const dim3 grid((p.view.W + BX - 1) / BX, (p.view.H + BY - 1) / BY, nz);
    composite_render_tma_kernel<TF, TMA_STAGES><<<grid, TMA_THREADS, smem, st>>>(P);
    return true;
}

}  // namespace vl3d
