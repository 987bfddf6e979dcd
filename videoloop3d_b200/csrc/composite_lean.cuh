// Instruction-lean composite kernels (included by composite.cu after CompositeParams / the v1 kernels).
//
// The v1 kernels were issue-bound (ncu: 64 % issue active, 27 % DRAM; ~480 SASS instructions per
// (pixel, plane, frame) sample in the backward).  These versions compute the same thing with about half
// the instructions:
//   * ex2.approx.ftz / rcp.approx.ftz (no denormal fix-up sequences around MUFU),
//   * 32-bit texel offsets against a per-frame 64-bit base held in registers: one IMAD.WIDE per tap
//     instead of a 4-instruction 64-bit add + shift chain,
//   * out-of-image threads replicate the border pixel, so every missing neighbour compares equal and
//     the eight per-direction pair masks collapse into two per-thread weights,
//   * neighbour exchange entirely through shared memory (1 STS.128 + 4 LDS.128 per frame; no shuffles),
//   * the shuffle hand-over of right-hand taps moves the 4-float sample gradient + two per-slot weights
//     instead of two 4-float products per frame (FFMA on the receiving side),
//   * static-tile work sits behind a warp-uniform branch,
//   * no per-frame "t < T" predicates: the host splits a launch into a TF-multiple and a TF=1 tail.
#pragma once

namespace vl3d {

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// same value as sigmoidf_fast (ex2.approx of -x*log2(e), rcp.approx) except that sub-1e-38 intermediates flush to 0
__device__ __forceinline__ float sigmoid_lean(float x) { return rcp_ftz(1.f + ex2_ftz(-1.4426950408889634f * x)); }

// The scratch grids are re-used every other plane while ~6 TB/s of Adam streams pass through L2: their lines are
// accessed with an evict_last policy so that the streams do not push them out to DRAM.
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void red_add_v4_hint(float4* addr, float4 v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ float4 ld_cg_hint(const float4* addr, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(addr), "l"(pol)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_cg_hint(float4* addr, float4 v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
template <typename P>
__device__ __forceinline__ P opaque_ptr(P p) {   // keep a base pointer as one 64-bit register pair
    asm volatile("" : "+l"(p));
    return p;
}

struct Geo {
    unsigned o00, o10, o01, o11;   // texel offsets inside one atlas frame
    float w00, w10, w01, w11;      // bilinear weights (0 for taps outside the atlas, MPV.py:425-427)
    int kind;                      // 0 none, 1 static, 2 dynamic
};

// quad-grid coordinates of pixel (u, v) on a plane; same arithmetic as plane_grid()
__device__ __forceinline__ bool plane_grid_lean(const float* __restrict__ h, float u, float v, float qwf, float qhf,
                                                float& gx, float& gy) {
    const float w = fmaf(h[6], u, fmaf(h[7], v, h[8]));
    float inv = rcp_ftz(w);
    inv = inv * fmaf(-w, inv, 2.f);
    gx = fmaf(h[0], u, fmaf(h[1], v, h[2])) * inv;
    gy = fmaf(h[3], u, fmaf(h[4], v, h[5])) * inv;
    return w > 0.f && gx > 0.f && gx < qwf && gy > 0.f && gy < qhf;
}

// quad under grid position (gx, gy) of plane d: returns its table entry (two 16-byte halves)
__device__ __forceinline__ const float4* quad_at(const CompositeParams& p, int d, float gx, float gy, int& qx, int& qy) {
    const int qw = p.view.qw, qh = p.view.qh;
    qx = min((int)gx, qw - 1); qy = min((int)gy, qh - 1);
    return reinterpret_cast<const float4*>(&p.quads[(d * qh + qy) * qw + qx]);
}

__device__ __forceinline__ Geo geo_from_quad(const CompositeParams& p, const float4* qp, const int4 qb, int qx, int qy,
                                             float gx, float gy) {
    Geo t;
    const float4 qa = __ldg(qp);
    const float a = gx - (float)qx, b = gy - (float)qy;
    const float lx = fmaf(a, qa.z, qa.x), ly = fmaf(b, qa.w, qa.y);
    const float flx = floorf(lx), fly = floorf(ly);
    const float fx = lx - flx, fy = ly - fly;
    const int ix = qb.x + (int)flx, iy = qb.y + (int)fly;            // >= 0: tiles lie inside the atlas (host-checked)
    t.kind = qb.z;
    const int aw = (t.kind == 2) ? p.view.dyn_w : p.view.sta_w;
    const int ah = (t.kind == 2) ? p.view.dyn_h : p.view.sta_h;
    // grid_sample zero padding can only trigger on the last row / column of the atlas
    const int cx0 = min(ix, aw - 1), cx1 = min(ix + 1, aw - 1);
    const int cy0 = min(iy, ah - 1), cy1 = min(iy + 1, ah - 1);
    const float gx1 = (ix + 1 < aw) ? fx : 0.f, gx0 = (ix < aw) ? 1.f - fx : 0.f;
    const float gy1 = (iy + 1 < ah) ? fy : 0.f, gy0 = (iy < ah) ? 1.f - fy : 0.f;
    const int r0 = cy0 * aw, r1 = cy1 * aw;
    t.o00 = (unsigned)(r0 + cx0); t.o10 = (unsigned)(r0 + cx1);
    t.o01 = (unsigned)(r1 + cx0); t.o11 = (unsigned)(r1 + cx1);
    t.w00 = gx0 * gy0; t.w10 = gx1 * gy0; t.w01 = gx0 * gy1; t.w11 = gx1 * gy1;
    return t;
}

__device__ __forceinline__ Geo geo_from_grid(const CompositeParams& p, int d, float gx, float gy) {
    int qx, qy;
    const float4* qp = quad_at(p, d, gx, gy, qx, qy);
    const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
    return geo_from_quad(p, qp, qb, qx, qy, gx, gy);
}

__device__ __forceinline__ float4 sample_lean(const float4* base, const Geo& t) {
    const float4 a = __ldg(base + t.o00), b = __ldg(base + t.o10), c = __ldg(base + t.o01), d = __ldg(base + t.o11);
    float4 r;
    r.x = a.x * t.w00 + b.x * t.w10 + c.x * t.w01 + d.x * t.w11;
    r.y = a.y * t.w00 + b.y * t.w10 + c.y * t.w01 + d.y * t.w11;
    r.z = a.z * t.w00 + b.z * t.w10 + c.z * t.w01 + d.z * t.w11;
    r.w = a.w * t.w00 + b.w * t.w10 + c.w * t.w01 + d.w * t.w11;
    r.x = sigmoid_lean(r.x); r.y = sigmoid_lean(r.y); r.z = sigmoid_lean(r.z); r.w = sigmoid_lean(r.w);
    return r;
}

// composite_lean's geo_from_quad plus the clamped integer tap coordinates (same arithmetic, same results)
struct GeoXY {
    Geo g;
    int cx0, cx1, cy0, cy1;
};

__device__ __forceinline__ GeoXY geoxy_from_quad(const CompositeParams& p, const float4* qp, const int4 qb, int qx, int qy,
                                                 float gx, float gy) {
    GeoXY r;
    const float4 qa = __ldg(qp);
    const float a = gx - (float)qx, b = gy - (float)qy;
    const float lx = fmaf(a, qa.z, qa.x), ly = fmaf(b, qa.w, qa.y);
    const float flx = floorf(lx), fly = floorf(ly);
    const float fx = lx - flx, fy = ly - fly;
    const int ix = qb.x + (int)flx, iy = qb.y + (int)fly;
    r.g.kind = qb.z;
    const int aw = (r.g.kind == 2) ? p.view.dyn_w : p.view.sta_w;
    const int ah = (r.g.kind == 2) ? p.view.dyn_h : p.view.sta_h;
    r.cx0 = min(ix, aw - 1); r.cx1 = min(ix + 1, aw - 1);
    r.cy0 = min(iy, ah - 1); r.cy1 = min(iy + 1, ah - 1);
    const float gx1 = (ix + 1 < aw) ? fx : 0.f, gx0 = (ix < aw) ? 1.f - fx : 0.f;
    const float gy1 = (iy + 1 < ah) ? fy : 0.f, gy0 = (iy < ah) ? 1.f - fy : 0.f;
    const int r0 = r.cy0 * aw, r1 = r.cy1 * aw;
    r.g.o00 = (unsigned)(r0 + r.cx0); r.g.o10 = (unsigned)(r0 + r.cx1);
    r.g.o01 = (unsigned)(r1 + r.cx0); r.g.o11 = (unsigned)(r1 + r.cx1);
    r.g.w00 = gx0 * gy0; r.g.w10 = gx1 * gy0; r.g.w01 = gx0 * gy1; r.g.w11 = gx1 * gy1;
    return r;
}

__device__ __forceinline__ float4 filter_taps(const float4 a, const float4 b, const float4 c, const float4 d, const Geo& t) {
    float4 r;
    r.x = a.x * t.w00 + b.x * t.w10 + c.x * t.w01 + d.x * t.w11;
    r.y = a.y * t.w00 + b.y * t.w10 + c.y * t.w01 + d.y * t.w11;
    r.z = a.z * t.w00 + b.z * t.w10 + c.z * t.w01 + d.z * t.w11;
    r.w = a.w * t.w00 + b.w * t.w10 + c.w * t.w01 + d.w * t.w11;
    r.x = sigmoid_lean(r.x); r.y = sigmoid_lean(r.y); r.z = sigmoid_lean(r.z); r.w = sigmoid_lean(r.w);
    return r;
}

// ------------------------------------------------------------------------------------------------
// pure render: planes in lockstep (CTA-uniform plane index => homography in uniform registers).
// Frames [tb, tb + TF*gridDim.z) of the call; the host guarantees they exist.
// ------------------------------------------------------------------------------------------------
template <int TF, int MINB>
__global__ void __launch_bounds__(BX* BY, MINB) composite_render_kernel(const __grid_constant__ CompositeParams p) {
    const int px = blockIdx.x * BX + threadIdx.x, py = blockIdx.y * BY + threadIdx.y;
    const int H = p.view.H, W = p.view.W;
    if (px >= W || py >= H) return;
    const int t0 = p.tb + blockIdx.z * TF;
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;
    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* fb[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int ft = p.ts ? __ldg(&p.ts[t0 + f]) : t0 + f;
        fb[f] = opaque_ptr(p.atlas_dyn + (size_t)ft * dyn_frame);
    }
    const float4* sb = opaque_ptr(p.atlas_sta);
    float Tr[TF], cr[TF], cg[TF], cb[TF], ca[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) { Tr[f] = 1.f; cr[f] = cg[f] = cb[f] = ca[f] = 0.f; }
    const int D = p.view.D;
    const float qwf = (float)p.view.qw, qhf = (float)p.view.qh;
    int nhit = 0;
    for (int d = 0; d < D; ++d) {
        float gx, gy;
        if (!plane_grid_lean(&p.view.hom[d * 9], u, v, qwf, qhf, gx, gy)) continue;
        const Geo tp = geo_from_grid(p, d, gx, gy);
        if (tp.kind == 0) continue;
        ++nhit;
        float4 val[TF];
        if (tp.kind == 2) {
#pragma unroll
            for (int f = 0; f < TF; ++f) val[f] = sample_lean(fb[f], tp);
        } else {
            const float4 s = sample_lean(sb, tp);                 // static tile: same for all frames (MPV.py:445)
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

// ------------------------------------------------------------------------------------------------
// backward (see the derivation above composite_bwd_v1_kernel).  Threads outside the image replicate the
// border pixel: a replica's value equals its in-image neighbour's bit for bit, so |a-b| and sign(a-b)
// vanish for every pair that does not exist; replicas never write.  What is left of the pair
// bookkeeping: horizontal pairs of the halo row and vertical pairs of the halo column belong to the
// neighbouring CTA (weights wH / wV zeroed), and the regulariser sums are kept by owned pixels only.
// ------------------------------------------------------------------------------------------------
// How the pixel rectangle [x0,x1] x [y0,y1] (inclusive) of a screen tile meets plane d: 0 = no pixel hits the
// plane, 1 = every pixel hits it, 2 = mixed / undecided.  The plane-grid coordinates are a projective image of
// the pixel rectangle, i.e. a convex quadrilateral: all four corners inside the plane rectangle => every pixel
// inside; bounding box outside on one side => nobody inside.  (Only meaningful when every quad of the plane
// exists: VL3D_VIEW_RECT_PLANES.)  `box` receives the top-left texel of the tile's atlas footprint (.x, .y) and its extent (.z, .w).
__device__ __forceinline__ int tile_plane_class(const CompositeParams& p, int d, int x0, int x1, int y0, int y1, int4& box) {
    const float qwf = (float)p.view.qw, qhf = (float)p.view.qh;
    const float cu[2] = {(float)x0 + 0.5f - p.view.cx, (float)x1 + 0.5f - p.view.cx};
    const float cv[2] = {(float)y0 + 0.5f - p.view.cy, (float)y1 + 0.5f - p.view.cy};
    const float* h = &p.view.hom[d * 9];
    float lxmin = 3e38f, lymin = 3e38f, lxmax = -3e38f, lymax = -3e38f;
    float gxmin = 3e38f, gxmax = -3e38f, gymin = 3e38f, gymax = -3e38f;
    int nfront = 0, ninside = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float gx, gy;
        const float u = cu[c & 1], v = cv[c >> 1];
        // (a margin of 1e-3 quad keeps the verdict independent of the last-bit rounding of interior pixels)
        const bool hit = plane_grid_lean(h, u, v, qwf, qhf, gx, gy);
        ninside += (hit && gx > 1e-3f && gx < qwf - 1e-3f && gy > 1e-3f && gy < qhf - 1e-3f) ? 1 : 0;
        nfront += (fmaf(h[6], u, fmaf(h[7], v, h[8])) > 0.f) ? 1 : 0;
        gxmin = fminf(gxmin, gx); gxmax = fmaxf(gxmax, gx); gymin = fminf(gymin, gy); gymax = fmaxf(gymax, gy);
        const float gxc = fminf(fmaxf(gx, 0.f), qwf), gyc = fminf(fmaxf(gy, 0.f), qhf);
        const int qx = min((int)gxc, p.view.qw - 1), qy = min((int)gyc, p.view.qh - 1);
        const float4* qp = reinterpret_cast<const float4*>(&p.quads[(d * p.view.qh + qy) * p.view.qw + qx]);
        const float4 qa = __ldg(qp);
        const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
        const float lx = (float)qb.x + fmaf(gxc - (float)qx, qa.z, qa.x), ly = (float)qb.y + fmaf(gyc - (float)qy, qa.w, qa.y);
        lxmin = fminf(lxmin, lx); lxmax = fmaxf(lxmax, lx);
        lymin = fminf(lymin, ly); lymax = fmaxf(lymax, ly);
    }
    // .z / .w: largest tap column / row of the footprint relative to the box origin (with a little slack for the
    // rounding of interior pixels): the footprint fits a TMA_BW x TMA_BH box iff .z < TMA_BW and .w < TMA_BH
    box = make_int4((int)floorf(lxmin), (int)floorf(lymin), 0, 0);
    box.z = (int)floorf(lxmax + 0.01f) + 1 - box.x;
    box.w = (int)floorf(lymax + 0.01f) + 1 - box.y;
    if (ninside == 4) return 1;
    if (nfront == 0) return 0;
    if (nfront == 4 && (gxmax <= -1e-3f || gxmin >= qwf + 1e-3f || gymax <= -1e-3f || gymin >= qhf + 1e-3f)) return 0;
    return 2;
}

__device__ __forceinline__ float sgn2(float c, float a, float b) {   // sign(c-a) + sign(c-b)
    return ((c > a ? 1.f : 0.f) - (c < a ? 1.f : 0.f)) + ((c > b ? 1.f : 0.f) - (c < b ? 1.f : 0.f));
}

// scatter one sample gradient `gl` (w.r.t. the pre-sigmoid bilinear sample) to its four taps at texel offsets
// q00 / q10 / q01 / q11 from `gb`.  Lane i's right-hand taps usually are lane i+1's left-hand taps: then the
// contribution travels by shuffle (gl_up with the sender's weights wu10 / wu11, zero if nothing is received) and is
// folded into the receiver's RED.
__device__ __forceinline__ void scatter_taps(float4* gb, const Geo& tp, const unsigned q00, const unsigned q10, const unsigned q01,
                                             const unsigned q11, const float4& gl, const float4& gl_up, float wu10, float wu11,
                                             bool sent0, bool sent1) {
    float4 l0, l1;
    l0.x = fmaf(gl_up.x, wu10, gl.x * tp.w00); l0.y = fmaf(gl_up.y, wu10, gl.y * tp.w00);
    l0.z = fmaf(gl_up.z, wu10, gl.z * tp.w00); l0.w = fmaf(gl_up.w, wu10, gl.w * tp.w00);
    l1.x = fmaf(gl_up.x, wu11, gl.x * tp.w01); l1.y = fmaf(gl_up.y, wu11, gl.y * tp.w01);
    l1.z = fmaf(gl_up.z, wu11, gl.z * tp.w01); l1.w = fmaf(gl_up.w, wu11, gl.w * tp.w01);
    red_add_v4(gb + q00, l0);
    red_add_v4(gb + q01, l1);
    if (!sent0) red_add_v4(gb + q10, make_float4(gl.x * tp.w10, gl.y * tp.w10, gl.z * tp.w10, gl.w * tp.w10));
    if (!sent1) red_add_v4(gb + q11, make_float4(gl.x * tp.w11, gl.y * tp.w11, gl.z * tp.w11, gl.w * tp.w11));
}
__device__ __forceinline__ void scatter_taps(float4* gb, const Geo& tp, const float4& gl, const float4& gl_up, float wu10,
                                             float wu11, bool sent0, bool sent1) {
    scatter_taps(gb, tp, tp.o00, tp.o10, tp.o01, tp.o11, gl, gl_up, wu10, wu11, sent0, sent1);
}
// owner mode: tap i goes to the scratch grid (`scr` + s_i, evict_last) if z_i, else to the gradient buffer (`gb` + o_i)
__device__ __forceinline__ void scatter_taps_own(float4* gb, float4* scr, const uint64_t pol, const Geo& tp, const unsigned zmask,
                                                 const int s00, const int s10, const int s01, const int s11, const float4& gl,
                                                 const float4& gl_up, float wu10, float wu11, bool sent0, bool sent1) {
    float4 l0, l1;
    l0.x = fmaf(gl_up.x, wu10, gl.x * tp.w00); l0.y = fmaf(gl_up.y, wu10, gl.y * tp.w00);
    l0.z = fmaf(gl_up.z, wu10, gl.z * tp.w00); l0.w = fmaf(gl_up.w, wu10, gl.w * tp.w00);
    l1.x = fmaf(gl_up.x, wu11, gl.x * tp.w01); l1.y = fmaf(gl_up.y, wu11, gl.y * tp.w01);
    l1.z = fmaf(gl_up.z, wu11, gl.z * tp.w01); l1.w = fmaf(gl_up.w, wu11, gl.w * tp.w01);
    if (zmask & 1u) red_add_v4_hint(scr + s00, l0, pol); else red_add_v4(gb + tp.o00, l0);
    if (zmask & 4u) red_add_v4_hint(scr + s01, l1, pol); else red_add_v4(gb + tp.o01, l1);
    if (!sent0) {
        const float4 r = make_float4(gl.x * tp.w10, gl.y * tp.w10, gl.z * tp.w10, gl.w * tp.w10);
        if (zmask & 2u) red_add_v4_hint(scr + s10, r, pol); else red_add_v4(gb + tp.o10, r);
    }
    if (!sent1) {
        const float4 r = make_float4(gl.x * tp.w11, gl.y * tp.w11, gl.z * tp.w11, gl.w * tp.w11);
        if (zmask & 8u) red_add_v4_hint(scr + s11, r, pol); else red_add_v4(gb + tp.o11, r);
    }
}

__device__ __forceinline__ float4 shfl_up4(const float4& v) {
    float4 r;
    r.x = __shfl_up_sync(0xffffffffu, v.x, 1); r.y = __shfl_up_sync(0xffffffffu, v.y, 1);
    r.z = __shfl_up_sync(0xffffffffu, v.z, 1); r.w = __shfl_up_sync(0xffffffffu, v.w, 1);
    return r;
}

// ------------------------------------------------------------------------------------------------
// "Owner" mode of the fused backward + Adam kernel (dense layout, regulariser tiling).
//
// A texel whose bilinear footprint is met by pixels of ONE screen tile only (and by no pixel that a neighbouring tile
// also processes as its halo) receives its complete gradient inside that tile.  Such texels never touch the gradient
// buffer in HBM: the tile accumulates a plane's texel gradients in a small per-CTA scratch box (RED.128 into a
// 30 KB region that lives in L2), and one slot later — behind the exchange barrier that exists anyway — reads the box
// back, runs Adam on the texels it owns (p, m, v read and written once, exclusively) and forwards the rest (texels
// near the tile border, shared with neighbours) to the gradient buffer as before, where the queue's ADAM items pick
// them up.  Ownership is a pure function of (plane, texel): the texel centre is mapped back to the screen with the
// plane's inverse homography; it is owned by tile X iff that point lies at least L pixels inside the pixels X alone
// processes, L >= the largest screen distance between a pixel and a texel it taps (host-computed bound per plane).
// The tile flush and the ADAM items evaluate the same predicate with the same arithmetic (no contraction: explicit
// fmaf / __fmul_rn / __fadd_rn), so every texel is updated exactly once.
// ------------------------------------------------------------------------------------------------
constexpr int OWN_ZW = 36, OWN_ZH = 7;                              // grid of candidate texels a tile may own on a plane (<= one per thread)
constexpr int OWN_ZN = OWN_ZW * OWN_ZH;                             // texels of one scratch grid
static_assert(OWN_ZN <= BX * BY, "one zone texel per thread");
static_assert(OWN_ZW >= BX && OWN_ZH <= BY - 1 && (OWN_ZW - BX) * OWN_ZH <= BX, "thread -> zone texel mapping of own_flush");

struct AdamK {
    float b1, b2, step_size, inv_sqrt_bc2, eps;
};
struct OwnParams {
    float hinv[VL3D_MAX_PLANES * 9];   // plane-local texel (x - rect.x, y - rect.y) -> continuous pixel index (pixel p's centre = p)
    float L[VL3D_MAX_PLANES];          // reach bound in pixels (see above); huge = nothing owned on this plane
    int4 rect[VL3D_MAX_PLANES];        // texels [x, z] x [y, w] (inclusive) that only this plane taps
    float4* scratch;                   // per CTA: [2][TF][OWN_ZN] texel gradients, all-zero between tiles
    const unsigned* table;             // per screen tile: planes it owns (own_table_kernel)
    const int2* zone;                  // per (screen tile, plane): top-left texel of the OWN_ZW x OWN_ZH grid that covers the
                                       // texels the tile owns on the plane (own_table_kernel)
    float4* m;
    float4* v;
    AdamK k;
    int gx, gy;                        // screen tiles
};

// same arithmetic as adam_kernel (optim.cu)
__device__ __forceinline__ void adam4(float4& pp, const float4 gg, float4& mm, float4& vv, const AdamK& K) {
    adam1(pp.x, gg.x, mm.x, vv.x, K.b1, K.b2, K.step_size, K.inv_sqrt_bc2, K.eps);
    adam1(pp.y, gg.y, mm.y, vv.y, K.b1, K.b2, K.step_size, K.inv_sqrt_bc2, K.eps);
    adam1(pp.z, gg.z, mm.z, vv.z, K.b1, K.b2, K.step_size, K.inv_sqrt_bc2, K.eps);
    adam1(pp.w, gg.w, mm.w, vv.w, K.b1, K.b2, K.step_size, K.inv_sqrt_bc2, K.eps);
}

__device__ __forceinline__ bool texel_to_pixel(const float* __restrict__ h, float x, float y, float& qx, float& qy) {
    const float w = fmaf(h[6], x, fmaf(h[7], y, h[8]));
    float inv = rcp_ftz(w);
    inv = __fmul_rn(inv, fmaf(-w, inv, 2.f));
    qx = __fmul_rn(fmaf(h[0], x, fmaf(h[1], y, h[2])), inv);
    qy = __fmul_rn(fmaf(h[3], x, fmaf(h[4], y, h[5])), inv);
    return w > 0.f;
}
// q at least L inside the pixels [lo + 1, lo + n - 2] that a tile starting at pixel `lo` (n threads) processes alone
__device__ __forceinline__ bool own_zone(float q, float L, int lo, int n) {
    return (__fadd_rn((float)lo, L) <= q) && (__fadd_rn(q, L) <= (float)(lo + n - 1));
}
// is texel (x, y) of plane d owned by screen tile (bx, by)?
__device__ __forceinline__ bool own_texel_of_tile(const OwnParams& O, int d, int x, int y, int bx, int by) {
    const int4 rc = O.rect[d];
    if (x < rc.x || x > rc.z || y < rc.y || y > rc.w) return false;
    float qx, qy;
    if (!texel_to_pixel(&O.hinv[d * 9], (float)(x - rc.x), (float)(y - rc.y), qx, qy)) return false;
    const float L = O.L[d];
    return own_zone(qx, L, bx * (BX - 1), BX) && own_zone(qy, L, by * (BY - 1), BY);
}
// is texel (x, y) of plane d owned by any screen tile?  (what the queue's ADAM items skip)
__device__ __forceinline__ bool own_texel_of_any(const OwnParams& O, int d, int x, int y) {
    const int4 rc = O.rect[d];
    if (x < rc.x || x > rc.z || y < rc.y || y > rc.w) return false;
    float qx, qy;
    if (!texel_to_pixel(&O.hinv[d * 9], (float)(x - rc.x), (float)(y - rc.y), qx, qy)) return false;
    const float L = O.L[d];
    if (!(L < 16.f) || !(qx > -64.f) || !(qy > -64.f) || !(qx < 1e6f) || !(qy < 1e6f)) return false;
    const int cx = (int)floorf((qx - L) * (1.f / (BX - 1))), cy = (int)floorf((qy - L) * (1.f / (BY - 1)));
    int bx = -1, by = -1;
#pragma unroll
    for (int c = 0; c < 2; ++c) {                                   // (the division may land one tile off)
        if (cx + c >= 0 && cx + c < O.gx && own_zone(qx, L, (cx + c) * (BX - 1), BX)) bx = cx + c;
        if (cy + c >= 0 && cy + c < O.gy && own_zone(qy, L, (cy + c) * (BY - 1), BY)) by = cy + c;
    }
    if (bx < 0 || by < 0) return false;
    return (__ldg(&O.table[by * O.gx + bx]) >> d) & 1u;
}

// Does screen tile (bx, by) own texels of plane d, and where?  Requires: every pixel of the tile hits the plane (cls == 1,
// from tile_plane_class, whose `box` is passed in), the footprint fits the scratch box, and the texels that can satisfy
// the ownership predicate — the image of the pixel rectangle [X0 + L, X1 - L] x [Y0 + L, Y1 - L] — fit a grid of
// OWN_ZW x OWN_ZH texels.  Returns the grid's top-left texel, or x = -1.  Evaluated ONCE per (tile, plane) by
// own_table_kernel; the tiles and the ADAM items read its verdict from memory, so there is one opinion only.
__device__ __forceinline__ int2 tile_own_zone(const CompositeParams& p, const OwnParams& O, int d, int bx, int by, int cls,
                                              const int4 box) {
    const int2 none = make_int2(-1, -1);
    if (cls != 1 || box.z >= TMA_BW || box.w >= TMA_BH) return none;
    const float L = O.L[d];
    if (!(L < 16.f)) return none;
    const float xa = (float)(bx * (BX - 1)) + L, xb = (float)(bx * (BX - 1) + BX - 1) - L;
    const float ya = (float)(by * (BY - 1)) + L, yb = (float)(by * (BY - 1) + BY - 1) - L;
    if (!(xa <= xb) || !(ya <= yb)) return none;
    const float qwf = (float)p.view.qw, qhf = (float)p.view.qh;
    const float cu[2] = {xa + 0.5f - p.view.cx, xb + 0.5f - p.view.cx};
    const float cv[2] = {ya + 0.5f - p.view.cy, yb + 0.5f - p.view.cy};
    const float* h = &p.view.hom[d * 9];
    float lxmin = 3e38f, lymin = 3e38f, lxmax = -3e38f, lymax = -3e38f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float gx, gy;
        plane_grid_lean(h, cu[c & 1], cv[c >> 1], qwf, qhf, gx, gy);
        const float gxc = fminf(fmaxf(gx, 0.f), qwf), gyc = fminf(fmaxf(gy, 0.f), qhf);
        const int qx = min((int)gxc, p.view.qw - 1), qy = min((int)gyc, p.view.qh - 1);
        const float4* qp = reinterpret_cast<const float4*>(&p.quads[(d * p.view.qh + qy) * p.view.qw + qx]);
        const float4 qa = __ldg(qp);
        const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
        const float lx = (float)qb.x + fmaf(gxc - (float)qx, qa.z, qa.x), ly = (float)qb.y + fmaf(gyc - (float)qy, qa.w, qa.y);
        lxmin = fminf(lxmin, lx); lxmax = fmaxf(lxmax, lx);
        lymin = fminf(lymin, ly); lymax = fmaxf(lymax, ly);
    }
    // texel centres inside the (convex) image of the rectangle, with slack for the rounding of the two maps
    const int tx0 = (int)ceilf(lxmin - 0.02f), tx1 = (int)floorf(lxmax + 0.02f);
    const int ty0 = (int)ceilf(lymin - 0.02f), ty1 = (int)floorf(lymax + 0.02f);
    if (tx1 - tx0 + 1 > OWN_ZW || ty1 - ty0 + 1 > OWN_ZH) return none;
    return make_int2(tx0, ty0);
}

// zone-grid texel of a thread: rows 0..OWN_ZH-1 x columns 0..BX-1 go to the warps in order (a row of the tile's threads =
// a row of texels), the remaining OWN_ZW - BX columns to the last warp.  Warps whose rows hold no owned texel skip Adam.
__device__ __forceinline__ bool own_zone_texel(int tx, int ty, int& col, int& row) {
    if (ty < OWN_ZH) { col = tx; row = ty; return true; }
    col = BX + tx % (OWN_ZW - BX); row = tx / (OWN_ZW - BX);
    return row < OWN_ZH;
}

// the flush of one plane's scratch grid (see above); every thread of the CTA calls it with its zone-grid texel.  All
// loads are issued before the first use, so a plane costs one memory round trip.
// Issued one slot ahead of own_flush: does this thread's zone texel of plane dd belong to the tile?  If so, pull its Adam
// state towards L2 so that the flush's loads do not wait for DRAM.
template <int TF>
__device__ __forceinline__ bool own_prefetch(const CompositeParams& p, const OwnParams& O, const int dd, const int2 zone, const int bx,
                                             const int by, const int t0) {
    int col, row;
    if (!own_zone_texel(threadIdx.x, threadIdx.y, col, row)) return false;
    const int x = zone.x + col, y = zone.y + row;
    if (!own_texel_of_tile(O, dd, x, y, bx, by)) return false;
    const int aw = p.view.dyn_w;
    const size_t frame = (size_t)p.view.dyn_h * aw;
    const size_t off = (size_t)t0 * frame + (size_t)y * aw + x;
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(O.m + off + f * frame));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(O.v + off + f * frame));
    }
    return true;
}

template <int TF>
__device__ __forceinline__ void own_flush(const CompositeParams& p, const OwnParams& O, const int2 zone, float4* scr, const bool mine,
                                          const int t0, const uint64_t pol) {
    int col, row;
    if (!own_zone_texel(threadIdx.x, threadIdx.y, col, row)) return;
    const int aw = p.view.dyn_w;
    const size_t frame = (size_t)p.view.dyn_h * aw;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int x = zone.x + col, y = zone.y + row;
    const int j = row * OWN_ZW + col;
    const size_t off = (size_t)t0 * frame + (size_t)y * aw + x;
    float4 g[TF], pp[TF], mm[TF], vv[TF];
#pragma unroll
    for (int f = 0; f < TF; ++f) g[f] = ld_cg_hint(scr + f * OWN_ZN + j, pol);
    if (mine) {
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            pp[f] = __ldcg(p.atlas_dyn + off + f * frame);           // (L2: the box of this plane was fetched a moment ago)
            mm[f] = __ldcs(O.m + off + f * frame);
            vv[f] = __ldcs(O.v + off + f * frame);
        }
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            adam4(pp[f], g[f], mm[f], vv[f], O.k);
            __stcs(const_cast<float4*>(p.atlas_dyn) + off + f * frame, pp[f]);
            __stcs(O.m + off + f * frame, mm[f]);
            __stcs(O.v + off + f * frame, vv[f]);
            st_cg_hint(scr + f * OWN_ZN + j, zero4, pol);
        }
    } else {
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            if (g[f].x != 0.f || g[f].y != 0.f || g[f].z != 0.f || g[f].w != 0.f) {
                red_add_v4(p.grad_dyn + off + f * frame, g[f]);
                st_cg_hint(scr + f * OWN_ZN + j, zero4, pol);
            }
        }
    }
}

// MODE 0: per-thread loads for every tile.  MODE 3 (VL3D_VIEW_RECT_PLANES): each tile decides at run time — if
// all its pixels hit the same planes (slot k == k-th plane for every pixel, so the planes can be walked in
// lockstep) it stages each plane's atlas footprint with TMA exactly like composite_render_tma_kernel: thread 0
// issues the box of plane k+NST-1 right after the exchange barrier of slot k, which is also what frees that stage;
// the remaining (image-border) tiles use the per-thread loads.
//
// bwd_tile = one (screen tile, chunk of TF frames) of the backward, as a device function: composite_bwd_kernel runs
// it once per CTA, the persistent fused backward + Adam kernel (fused_bwd_adam.cu) once per work item.  `kbase`
// counts the TMA stage uses of this CTA so far (the mbarrier phases carry over from tile to tile); `first` = the
// CTA's first tile (initialises the mbarriers).  Callers separate two tiles by a __syncthreads().
// stages of the box ring (boxes in flight = stages - 1).  Two are enough and leave 15 KB per CTA to L1: fused pass
// 40.1 vs 40.7 ms at 720p, 2.96 vs 3.04 ms at 180x320, standalone backward unchanged (18.6 ms).
#ifndef VL3D_BWD_STAGES
#define VL3D_BWD_STAGES 2
#endif
constexpr int BWD_TMA_STAGES = VL3D_BWD_STAGES;

template <int TF, bool SMOOTH, int MODE, bool OWN = false>
__device__ __forceinline__ void bwd_tile(const TmaRenderParams& P, const int bx, const int by, const int t0, unsigned& kbase,
                                         const bool first, const OwnParams* const O = nullptr) {
    const CompositeParams& p = P.p;
    static_assert(MODE == 0 || SMOOTH, "the split launch is only built for the regulariser tiling");
    static_assert(!OWN || MODE >= 2, "owner mode rides on the TMA-staged path");
    constexpr int SX = SMOOTH ? BX - 1 : BX, SY = SMOOTH ? BY - 1 : BY;
    constexpr int NST = BWD_TMA_STAGES;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int px0 = bx * SX + tx, py0 = by * SY + ty;
    const int H = p.view.H, W = p.view.W;
    const bool active = px0 < W && py0 < H;
    const bool owned = active && tx < SX && ty < SY;
    const int px = min(px0, W - 1), py = min(py0, H - 1);          // replicas of the border pixel
    const float u = (float)px + 0.5f - p.view.cx, v = (float)py + 0.5f - p.view.cy;

    __shared__ float4 s_ex[SMOOTH ? 2 : 1][SMOOTH ? TF : 1][SMOOTH ? BY : 1][SMOOTH ? BX : 1];
    __shared__ int s_cls[3];                                        // (planes hit by every pixel, any mixed plane, footprint fits)
    __shared__ int4 s_box[MODE >= 2 ? VL3D_MAX_PLANES : 1];
    __shared__ int2 s_zone[OWN ? VL3D_MAX_PLANES : 1];
    __shared__ __align__(8) uint64_t s_full[MODE >= 2 ? NST : 1];
    extern __shared__ __align__(128) unsigned char bwd_dyn_smem[];  // MODE 2: [NST][TF][TMA_BH][TMA_BW] texels
    unsigned in_mask = 0u, own_mask = 0u;
    bool use_tma = false;
    if (MODE != 0) {
        if (ty == 0) {
            int cls = 0;
            int4 box = make_int4(0, 0, 0, 0);
            if (tx < p.view.D)
                cls = tile_plane_class(p, tx, bx * SX, min(bx * SX + BX - 1, W - 1), by * SY, min(by * SY + BY - 1, H - 1), box);
            const unsigned m_in = __ballot_sync(0xffffffffu, cls == 1), m_mixed = __ballot_sync(0xffffffffu, cls == 2);
            if (MODE >= 2 && tx < p.view.D) s_box[tx] = box;
            unsigned m_own = 0u;
            if (OWN) {                                              // own_table_kernel's verdict for this tile
                const int gxt = (W + SX - 1) / SX;
                int2 z = make_int2(-1, -1);
                if (tx < p.view.D) z = __ldg(&O->zone[(size_t)(by * gxt + bx) * VL3D_MAX_PLANES + tx]);
                s_zone[tx] = z;
                m_own = __ballot_sync(0xffffffffu, z.x >= 0);
            }
            if (tx == 0) {
                s_cls[0] = (int)m_in; s_cls[1] = (int)m_mixed; s_cls[2] = (int)m_own;
                if (MODE >= 2 && first) {
#pragma unroll
                    for (int s = 0; s < NST; ++s) mbar_init(&s_full[s], 1);
                    mbar_fence_init();
                }
            }
        }
        __syncthreads();
        const bool uniform = s_cls[1] == 0;
        in_mask = (unsigned)s_cls[0];
        use_tma = MODE >= 2 && uniform;
        if (OWN) own_mask = use_tma ? (unsigned)s_cls[2] & in_mask : 0u;
    }
    // owner mode: this CTA's scratch boxes, the plane flushed one slot behind
    float4* const scr = OWN ? O->scratch + (size_t)blockIdx.x * (2 * TF * OWN_ZN) : nullptr;
    const uint64_t pol = OWN ? l2_evict_last_policy() : 0ull;
    int prev_dd = -1;
    bool prev_mine = false;

    const size_t dyn_frame = (size_t)p.view.dyn_h * p.view.dyn_w;
    const float4* ab[TF];
    float4* gb[TF];
    float g0[TF], g1[TF], g2[TF], tot[TF], Tr[TF], pre[TF];
    const size_t plane = (size_t)H * W;
    const size_t pix = (size_t)py * W + px;
#pragma unroll
    for (int f = 0; f < TF; ++f) {
        const int t = t0 + f;
        const int ft = p.ts ? __ldg(&p.ts[t]) : t;
        ab[f] = opaque_ptr(p.atlas_dyn + (size_t)ft * dyn_frame);
        gb[f] = opaque_ptr(p.grad_dyn + (size_t)ft * dyn_frame);
        g0[f] = g1[f] = g2[f] = 0.f; tot[f] = 0.f; Tr[f] = 1.f; pre[f] = 0.f;
        if (owned) {
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
    const float4* sb = opaque_ptr(p.atlas_sta);
    float4* gsb = opaque_ptr(p.grad_sta);

    float wHc = 0.f, wVc = 0.f, wHa = 0.f, wVa = 0.f;
    // shared-memory neighbours, clamped to the tile (a clamped read returns the thread's own value)
    const int o_c = ty * BX + tx;
    const int o_r = ty * BX + min(tx + 1, BX - 1), o_l = ty * BX + max(tx - 1, 0);
    const int o_d = min(ty + 1, BY - 1) * BX + tx, o_u = max(ty - 1, 0) * BX + tx;
    if (SMOOTH) {
        if (ty < SY) { wHc = __ldg(p.w_smooth); wHa = __ldg(p.w_smooth + 2); }
        if (tx < SX) { wVc = __ldg(p.w_smooth + 1); wVa = __ldg(p.w_smooth + 3); }
    }
    const bool want_sums = SMOOTH && p.smooth != nullptr;
    float sxr = 0.f, syr = 0.f, sxa = 0.f, sya = 0.f;

    const int D = p.view.D;
    const float qwf = (float)p.view.qw, qhf = (float)p.view.qh;
    // per-thread path: which planes does this pixel's ray hit on an existing quad?  All D tests up front — their
    // quad-table loads are independent and overlap; walking the planes lazily inside the slot loop chained up to D
    // dependent loads per pixel and made the backward of a tile-culled model as slow as the dense one (18.4 ms at
    // 720p for a quarter of the samples).
    unsigned hits = 0u;
    if (!use_tma) {
#pragma unroll 8
        for (int dd = 0; dd < D; ++dd) {
            float gx, gy;
            int kind = 0;
            if (plane_grid_lean(&p.view.hom[dd * 9], u, v, qwf, qhf, gx, gy)) {
                int qx, qy;
                const float4* qp = quad_at(p, dd, gx, gy, qx, qy);
                kind = __ldg(&reinterpret_cast<const int4*>(qp + 1)->z);
            }
            hits |= (kind != 0 ? 1u : 0u) << dd;
        }
    }
    // MODE 2: planes of this tile in order, TMA issue state of thread 0
    const int nplanes = __popc(in_mask);
    unsigned rem_planes = in_mask, rem_issue = in_mask;
    unsigned k_issue = kbase;
    float4* tiles = reinterpret_cast<float4*>(bwd_dyn_smem);
    auto issue_plane = [&]() {                                      // thread 0 only
        const int di = __ffs(rem_issue) - 1;
        rem_issue &= rem_issue - 1u;
        const int s = (int)(k_issue % NST);
        const int4 bi = s_box[di];
        mbar_arrive_expect_tx(&s_full[s], TF * TMA_BOX_BYTES);
#pragma unroll
        for (int f = 0; f < TF; ++f)
            tma_load_3d(tiles + (size_t)(s * TF + f) * (TMA_BOX_BYTES / 16), &P.tmap, &s_full[s], bi.x * 4, bi.y, t0 + f);
        ++k_issue;
    };
    if (use_tma && tx == 0 && ty == 0) {
#pragma unroll
        for (int i = 0; i < NST - 1; ++i)
            if (nplanes > i) issue_plane();
    }
    for (int k = 0;; ++k) {
        Geo tp;
        tp.kind = 0; tp.o00 = tp.o10 = tp.o01 = tp.o11 = 0u;
        tp.w00 = tp.w10 = tp.w01 = tp.w11 = 0.f;
        float4 val[TF];
#pragma unroll
        for (int f = 0; f < TF; ++f) val[f] = make_float4(0.f, 0.f, 0.f, 0.f);   // zero canvas (MPV.py:441)
        unsigned zmask = 0u;                                        // owner mode: which taps of this sample go to the scratch grid
        int s00 = 0, s10 = 0, s01 = 0, s11 = 0;
        int cur_dd = -1;
        if (MODE >= 2 && use_tma) {
            // every pixel of the tile (replicas included) hits exactly the planes of in_mask: slot k = k-th plane
            if (k >= nplanes) break;
            const int dd = __ffs(rem_planes) - 1;
            cur_dd = dd;
            rem_planes &= rem_planes - 1u;
            const unsigned use = kbase + (unsigned)k;
            const int s = (int)(use % NST);
            float gx, gy;
            plane_grid_lean(&p.view.hom[dd * 9], u, v, qwf, qhf, gx, gy);
            int qx, qy;
            const float4* qp = quad_at(p, dd, gx, gy, qx, qy);
            const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
            const GeoXY t = geoxy_from_quad(p, qp, qb, qx, qy, gx, gy);
            tp = t.g;
            const int4 bi = s_box[dd];
            const int lx0 = t.cx0 - bi.x, lx1 = t.cx1 - bi.x, ly0 = t.cy0 - bi.y, ly1 = t.cy1 - bi.y;
            if (OWN && ((own_mask >> dd) & 1u)) {                   // taps inside the plane's zone grid accumulate in the scratch grid
                const int2 z = s_zone[dd];
                const int zx0 = t.cx0 - z.x, zx1 = t.cx1 - z.x, zy0 = t.cy0 - z.y, zy1 = t.cy1 - z.y;
                const bool x0in = (unsigned)zx0 < (unsigned)OWN_ZW, x1in = (unsigned)zx1 < (unsigned)OWN_ZW;
                const bool y0in = (unsigned)zy0 < (unsigned)OWN_ZH, y1in = (unsigned)zy1 < (unsigned)OWN_ZH;
                zmask = (x0in && y0in ? 1u : 0u) | (x1in && y0in ? 2u : 0u) | (x0in && y1in ? 4u : 0u) | (x1in && y1in ? 8u : 0u);
                s00 = zy0 * OWN_ZW + zx0; s10 = zy0 * OWN_ZW + zx1; s01 = zy1 * OWN_ZW + zx0; s11 = zy1 * OWN_ZW + zx1;
            }
            mbar_wait(&s_full[s], (use / NST) & 1u);                // the plane's boxes have landed
            if (tp.kind == 2 && lx0 >= 0 && lx1 < TMA_BW && ly0 >= 0 && ly1 < TMA_BH) {
                const float4* tb = tiles + (size_t)(s * TF) * (TMA_BOX_BYTES / 16);
                const int a00 = ly0 * TMA_BW + lx0, a10 = ly0 * TMA_BW + lx1, a01 = ly1 * TMA_BW + lx0, a11 = ly1 * TMA_BW + lx1;
#pragma unroll
                for (int f = 0; f < TF; ++f) {
                    const float4* tf = tb + f * (TMA_BOX_BYTES / 16);
                    val[f] = filter_taps(tf[a00], tf[a10], tf[a01], tf[a11], tp);
                }

            } else if (tp.kind == 2) {
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = sample_lean(ab[f], tp);
            } else if (tp.kind == 1) {
                const float4 sv = sample_lean(sb, tp);
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = sv;
            }
        } else {
            // slot k of this pixel = its k-th hit plane along the ray (utils.py:64-69): the next set bit of the hit mask
            // built in the prologue.  The loop ends when no thread of the tile (warp, without the regulariser) has a
            // slot left.
            if (hits != 0u) {
                const int dd = __ffs(hits) - 1;
                hits &= hits - 1u;
                float gx, gy;
                plane_grid_lean(&p.view.hom[dd * 9], u, v, qwf, qhf, gx, gy);
                int qx, qy;
                const float4* qp = quad_at(p, dd, gx, gy, qx, qy);
                const int4 qb = __ldg(reinterpret_cast<const int4*>(qp + 1));
                tp = geo_from_quad(p, qp, qb, qx, qy, gx, gy);
            }
            if (tp.kind == 2) {
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = sample_lean(ab[f], tp);
            } else if (tp.kind == 1) {
                const float4 sv = sample_lean(sb, tp);
#pragma unroll
                for (int f = 0; f < TF; ++f) val[f] = sv;
            }
        }
        const bool has = tp.kind != 0;
        if (!SMOOTH) {
            if (!__any_sync(0xffffffffu, has)) break;
        }
        float4 gs[TF];   // dL/d(activated value) from the smoothness terms
#pragma unroll
        for (int f = 0; f < TF; ++f) gs[f] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SMOOTH) {
            float4* ex = &s_ex[k & 1][0][0][0];
#pragma unroll
            for (int f = 0; f < TF; ++f) ex[f * (BX * BY) + o_c] = val[f];
            if (MODE >= 2 && use_tma) {
                __syncthreads();                                    // also: everybody is done with the stage of slot k-1
                if (tx == 0 && ty == 0 && k + NST - 1 < nplanes) issue_plane();
            } else if (!__syncthreads_or(has)) break;              // (block-uniform)
#pragma unroll
            for (int f = 0; f < TF; ++f) {
                const float4 c = val[f];
                const float4 r = ex[f * (BX * BY) + o_r], l = ex[f * (BX * BY) + o_l];
                const float4 dn = ex[f * (BX * BY) + o_d], up = ex[f * (BX * BY) + o_u];
                // d|a-b|/da = sign(a-b); pairs that do not exist compare a value with itself
                gs[f].x = fmaf(wHc, sgn2(c.x, r.x, l.x), wVc * sgn2(c.x, dn.x, up.x));
                gs[f].y = fmaf(wHc, sgn2(c.y, r.y, l.y), wVc * sgn2(c.y, dn.y, up.y));
                gs[f].z = fmaf(wHc, sgn2(c.z, r.z, l.z), wVc * sgn2(c.z, dn.z, up.z));
                gs[f].w = fmaf(wHa, sgn2(c.w, r.w, l.w), wVa * sgn2(c.w, dn.w, up.w));
                if (want_sums) {                                   // the regulariser values themselves (MPV.py:517-531)
                    sxr += fabsf(c.x - r.x) + fabsf(c.y - r.y) + fabsf(c.z - r.z);
                    sxa += fabsf(c.w - r.w);
                    syr += fabsf(c.x - dn.x) + fabsf(c.y - dn.y) + fabsf(c.z - dn.z);
                    sya += fabsf(c.w - dn.w);
                }
            }
            // the other buffer is rewritten next iteration; its readers finished before this barrier
        }
        // ---- hand-over keys: (kind, texel offset) of the taps a lane writes / could absorb; lanes that do not
        // write (replicas, empty slots) carry keys that match nothing
        const bool wr = active && has;
        const unsigned ktag = (unsigned)tp.kind << 30;
        const unsigned rk0 = wr ? (tp.o00 | ktag) : 0xffffffffu, rk1 = wr ? (tp.o01 | ktag) : 0xffffffffu;
        const unsigned sk0 = wr ? (tp.o10 | ktag) : 0xfffffffeu, sk1 = wr ? (tp.o11 | ktag) : 0xfffffffeu;
        const unsigned up_sk0 = __shfl_up_sync(0xffffffffu, sk0, 1), up_sk1 = __shfl_up_sync(0xffffffffu, sk1, 1);
        const unsigned dn_rk0 = __shfl_down_sync(0xffffffffu, rk0, 1), dn_rk1 = __shfl_down_sync(0xffffffffu, rk1, 1);
        const bool recv0 = (tx > 0) & (up_sk0 == rk0), recv1 = (tx > 0) & (up_sk1 == rk1);
        const bool sent0 = (tx < BX - 1) & (dn_rk0 == sk0), sent1 = (tx < BX - 1) & (dn_rk1 == sk1);
        const float wu10s = __shfl_up_sync(0xffffffffu, tp.w10, 1), wu11s = __shfl_up_sync(0xffffffffu, tp.w11, 1);
        const float wu10 = recv0 ? wu10s : 0.f, wu11 = recv1 ? wu11s : 0.f;
        float4 gsta = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int f = 0; f < TF; ++f) {
            // gradient w.r.t. the pre-sigmoid bilinear sample; an empty slot has c = 0, hence gl = 0
            const float4 c = val[f];
            const float a = c.w, om = 1.f - a;
            const float bw = a * Tr[f];
            const float gc = g0[f] * c.x + g1[f] * c.y + g2[f] * c.z;
            pre[f] = fmaf(bw, gc, pre[f]);
            const float S = tot[f] - pre[f];
            float4 gl;
            gl.x = fmaf(g0[f], bw, gs[f].x) * (c.x - c.x * c.x);
            gl.y = fmaf(g1[f], bw, gs[f].y) * (c.y - c.y * c.y);
            gl.z = fmaf(g2[f], bw, gs[f].z) * (c.z - c.z * c.z);
            gl.w = a * (om * fmaf(Tr[f], gc, gs[f].w) - S);
            Tr[f] *= om;
            gsta.x += gl.x; gsta.y += gl.y; gsta.z += gl.z; gsta.w += gl.w;   // static tiles: sum over frames (MPV.py:445)
            const float4 gl_up = shfl_up4(gl);
            if (wr && tp.kind == 2) {
                if (OWN && zmask != 0u)
                    scatter_taps_own(gb[f], scr + ((k & 1) * TF + f) * OWN_ZN, pol, tp, zmask, s00, s10, s01, s11, gl, gl_up, wu10,
                                     wu11, sent0, sent1);
                else
                    scatter_taps(gb[f], tp, gl, gl_up, wu10, wu11, sent0, sent1);
            }
        }
        if (__any_sync(0xffffffffu, tp.kind == 1)) {
            const float4 gl_up = shfl_up4(gsta);
            if (wr && tp.kind == 1) scatter_taps(gsb, tp, gsta, gl_up, wu10, wu11, sent0, sent1);
        }
        if (OWN && use_tma) {
            // the previous slot's scratch box is complete (its REDs precede this slot's exchange barrier): Adam on the
            // texels this tile owns, the rest on to the gradient buffer; the box is zero again before slot k + 1 reuses it
            const bool cur_own = (own_mask >> cur_dd) & 1u;
            const bool cur_mine = cur_own && own_prefetch<TF>(p, *O, cur_dd, s_zone[cur_dd], bx, by, t0);
            if (prev_dd >= 0) own_flush<TF>(p, *O, s_zone[prev_dd], scr + (((k - 1) & 1) * TF) * OWN_ZN, prev_mine, t0, pol);
            prev_dd = cur_own ? cur_dd : -1;
            prev_mine = cur_mine;
        }
    }
    if (OWN && use_tma && prev_dd >= 0) {
        __syncthreads();                                            // the last slot's REDs
        own_flush<TF>(p, *O, s_zone[prev_dd], scr + (((nplanes - 1) & 1) * TF) * OWN_ZN, prev_mine, t0, pol);
    }
    if (want_sums) {
        __shared__ float s_sum[4][(BX * BY) / 32];
        const int warp = (ty * BX + tx) >> 5;
        const float m = owned ? 1.f : 0.f;
        const float a0 = warp_sum(sxr * m), a1 = warp_sum(syr * m), a2 = warp_sum(sxa * m), a3 = warp_sum(sya * m);
        if (tx == 0) { s_sum[0][warp] = a0; s_sum[1][warp] = a1; s_sum[2][warp] = a2; s_sum[3][warp] = a3; }
        __syncthreads();
        if (ty == 0 && tx < 4) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < (BX * BY) / 32; ++i) acc += (double)s_sum[tx][i];
            atomicAdd(&p.smooth[tx], acc);
        }
    }
    if (use_tma) kbase += (unsigned)nplanes;
}

template <int TF, bool SMOOTH, int MODE>
__global__ void __launch_bounds__(BX* BY, TF <= 2 ? 3 : 2) composite_bwd_kernel(const __grid_constant__ TmaRenderParams P) {
    unsigned kbase = 0u;
    bwd_tile<TF, SMOOTH, MODE>(P, (int)blockIdx.x, (int)blockIdx.y, P.p.tb + (int)blockIdx.z * TF, kbase, true);
}

}  // namespace vl3d
