/*
 * vl3d.h — C ABI of libvl3d.so: the B200 (sm_100a) kernels behind VideoLoop3D's stage-2 hot path.
 *
 * The reference (limacv/VideoLoop3D) is pure Python and has no FFI / plugin registry: the boundary
 * it offers is the Python operator surface (`MPMeshVid.render` / `forward`, the loop-loss callables,
 * `run_iter`).  This header is the thin C boundary underneath our drop-in Python mirror of that
 * surface (`videoloop3d_b200/`); each entry point cites the reference code it replaces.
 * `INTEGRATION.md` shows the ctypes stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named `*_host` or it is
 *     a `const vl3d_*` descriptor struct (host memory, copied into the launch);
 *   - the caller allocates every input, output and workspace; the library never allocates, frees or
 *     keeps state between calls (except the last-error string);
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), no host synchronisation inside,
 *     CUDA-graph capturable;
 *   - return 0 on success, a negative `VL3D_E*` for argument errors (validated before any launch),
 *     a positive value = cudaError_t from the launch; `vl3d_last_error_string()` describes the last
 *     failure on the calling thread.  Never throws, never aborts.
 *   - fp32 arithmetic throughout (the reference's `fp16` flag is marked "do NOT use",
 *     config_parser.py:32-33); NN indices are int32.
 *   - texel layout: atlases are RGBA-interleaved, i.e. a torch tensor of logical shape (T,4,Hd,Wd)
 *     in `channels_last` memory format == physical (T,Hd,Wd,4).  Rendered video / targets are
 *     frame-major channel-planar (T,3,H,W).
 */
#ifndef VL3D_H_
#define VL3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VL3D_VERSION 100
#define VL3D_MAX_PLANES 32

#define VL3D_EINVAL   (-1)  /* bad argument / shape */
#define VL3D_ENULL    (-2)  /* required pointer is NULL */
#define VL3D_ERANGE   (-3)  /* size outside the supported range (e.g. D > 32, patch too large) */
#define VL3D_EALIGN   (-4)  /* pointer not 16-byte aligned where float4 access is required */

/* One quad ("tile") of the MPV mesh: where its texels live in the atlas.
 * Built on the host from faces(_dyn)[::2,0] and uvs(_dyn) (MPV.py:66-81, MPI.py:385-418).
 * Atlas pixel coordinate of a point with in-quad fractions (a, b):
 *     x = x0i + (x0f + a * sx),   y = y0i + (y0f + b * sy)      (align_corners=True pixels)
 * kind: 0 = culled / absent, 1 = static atlas, 2 = dynamic atlas. */
typedef struct vl3d_quad {
    float   x0f, y0f, sx, sy;
    int32_t x0i, y0i, kind, reserved;
} vl3d_quad;

/* Per-call view descriptor (host memory).  Replaces the no_grad geometry block of
 * MPMeshVid.render (MPV.py:353-405: NDC transform, pytorch3d rasterize_meshes, get_uvs):
 * for plane d, target pixel centre (c+0.5, r+0.5) maps to quad-grid coordinates
 *     [gx*w, gy*w, w] = hom[d] * [c + 0.5 - cx, r + 0.5 - cy, 1],
 * gx in (0,qw), gy in (0,qh); the homography is pre-multiplied so that a valid (in front of the
 * camera) intersection has w > 0.  Planes are ordered nearest first (MPV.py:51). */
typedef struct vl3d_view {
    int32_t H, W;            /* rendered patch size */
    int32_t D;               /* planes, 1..VL3D_MAX_PLANES */
    int32_t qh, qw;          /* quads per plane = (mpi_h_verts-1, mpi_w_verts-1) */
    int32_t dyn_h, dyn_w;    /* dynamic atlas size in texels */
    int32_t sta_h, sta_w;    /* static atlas size in texels */
    float   cx, cy;
    float   hom[VL3D_MAX_PLANES * 9];
    int32_t flags;           /* VL3D_VIEW_* bits (hints; results never depend on them) */
} vl3d_view;

/* vl3d_view.flags: every plane's quads are dynamic and tile ONE axis-aligned rectangle of atlas_dyn contiguously
 * (the dense layout of MPV.py:75-81), so the footprint of a screen tile on a plane is one atlas rectangle and the
 * render may fetch it with a single TMA box per (tile, plane, frame) instead of four loads per pixel. */
#define VL3D_VIEW_RECT_PLANES 1

/* Looping-loss problem descriptor (fitted crop; utils_vid.py:305-320).
 * x = rendered (looped, scaled) video, frames tx in [0, t); y = target video, frames in [0, F).
 * Both are channel-planar with explicit element strides (frame, channel, row); pixels contiguous. */
typedef struct vl3d_loss_desc {
    int32_t t, F;            /* frames used of x (fitted) and y */
    int32_t h, w;            /* fitted spatial crop */
    int32_t p, pt, s, st;    /* patch size, temporal patch size, stride, temporal stride */
    int32_t n1, n2;          /* query / candidate temporal positions */
    int32_t ho, wo;          /* patch positions per column / row */
    int64_t x_sf, x_sc, x_sr;/* x strides in elements: frame, channel, row */
    int64_t y_sf, y_sc, y_sr;
    int32_t use_alpha;       /* 0: alpha=None (alpha > 100, utils_vid.py:208) */
    float   alpha;
} vl3d_loss_desc;

int         vl3d_version(void);
const char* vl3d_last_error_string(void);

/* ---- composite (MPV.py:413-475 render_masked_rgba + masked_scatter + utils_mpi.overcompose,
 *      and the slot-wise smoothness regularisers MPV.py:517-531) -------------------------------
 * ts:        T int32 frame ids into atlas_dyn, or NULL for 0..T-1 (MPV.py:480-481).
 * rgb_out:   (T + pad, 3, H, W); frames t < pad are also written at T + t (loop pad, MPV.py:490-492).
 * alpha_out: (T, H, W) or NULL  (MPV.py:454).
 * smooth_sums: 4 doubles, ACCUMULATED: sum|dx rgb|, sum|dy rgb|, sum|dx a|, sum|dy a| over slot-indexed
 *            values (missing slots read as 0), or NULL to skip the regulariser pass.
 * mpi_out:   (T, H, W, D, 4) slot-indexed activated RGBA (variables['mpi'], MPV.py:441-449) or NULL.
 * hits_out:  (H, W) int32 number of occupied slots per pixel (K = max, utils.py:64-69) or NULL. */
int vl3d_composite_fwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                       const float* atlas_sta, const int32_t* ts, int32_t T, int32_t pad,
                       float* rgb_out, float* alpha_out, double* smooth_sums, float* mpi_out,
                       int32_t* hits_out, void* stream);

/* Hand-written backward of the above (replaces autograd through MPV.py:413-451, utils_mpi.py:92-107
 * and MPV.py:517-531).  grad_rgb: (T + pad, 3, H, W) = dL/d rgb_out (pad frames folded in-kernel).
 * rgb: the forward's rgb_out.  w_smooth: 4 DEVICE floats dL/d(smooth_sums[i]), or NULL when the forward
 * ran without the regulariser pass (no host round trip of upstream scalars).  smooth_sums: optional 4
 * doubles (ACCUMULATED) — the backward can emit the regulariser sums itself (it exchanges the same
 * neighbour values anyway), which lets a fused step run the forward as a pure render.
 * grad_dyn (Tall,Hd,Wd,4) and grad_sta (Hs,Ws,4) are ACCUMULATED into (caller zeroes them). */
int vl3d_composite_bwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                       const float* atlas_sta, const int32_t* ts, int32_t T, int32_t pad,
                       const float* grad_rgb, const float* rgb, const float* w_smooth,
                       double* smooth_sums, float* grad_dyn, float* grad_sta, void* stream);

/* ---- optional per-ray terms of the composite and their backward (off in every shipped stage-2 config) --------------------
 * Replaces: MPV.py:454 `alpha = blend_weight.sum(-1)` as a differentiable output (background blend MPV.py:455-461,
 * density regulariser MPV.py:533-536), MPV.py:384-385,463-464 `disp = (1 / zbuf * blend_weight).sum(-1)` (d_smooth,
 * MPV.py:538-551), the sparsity regulariser MPV.py:511-515 and autograd's backward through them; the stage-1 model uses
 * the same terms (MPI.py:552-566,603-607,647-650).
 * inv_depth_host: HOST array of 3*D floats (copied into the launch), per plane (a, b, c) with
 *     1 / (view depth of the ray-plane intersection) = a * (c + 0.5 - cx) + b * (r + 0.5 - cy) + c
 *     (exactly linear in the pixel for a plane; any affine map of it, e.g. MPI.py:553's normalisation, folds into a, b, c);
 *     NULL when neither disp_out nor grad_disp is used.
 * fwd: alpha_out (T,H,W) = sum_k bw_k, disp_out (T,H,W) = sum_k bw_k / depth_k, sparsity_sum[0] += sum over frames and
 *     pixels of |a|_1 / max(|a|_2, sparsity_eps) over the ray's slot alphas (each optional).
 * bwd: given dL/d alpha_out, dL/d disp_out (T,H,W each, optional) and w_sparsity = DEVICE float dL/d sparsity_sum
 *     (optional), ACCUMULATES the resulting gradients of the alpha logits into grad_dyn / grad_sta (same buffers and
 *     layout as vl3d_composite_bwd; these terms give no gradient to the colour channels). */
int vl3d_composite_terms_fwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                             const float* atlas_sta, const int32_t* ts, int32_t T, const float* inv_depth_host,
                             float sparsity_eps, float* alpha_out, float* disp_out, double* sparsity_sum, void* stream);
int vl3d_composite_terms_bwd(const vl3d_view* view, const vl3d_quad* quads, const float* atlas_dyn,
                             const float* atlas_sta, const int32_t* ts, int32_t T, const float* inv_depth_host,
                             float sparsity_eps, const float* grad_alpha, const float* grad_disp,
                             const float* w_sparsity, float* grad_dyn, float* grad_sta, void* stream);

/* ---- scale-invariant gain (MPV.py:499-504):
 *      out[0] = (exp(mean_{c,h,w} log((mean_F res + .01)/(mean_T rgb + .01))) + 3) / 4
 * rgb (T,3,H,W) contiguous, res (F,3,H,W) contiguous. partials: workspace of >= vl3d_scale_partials()
 * doubles. */
int vl3d_scale_partials(void);
int vl3d_scale_invariant(const float* rgb, int32_t T, const float* res, int32_t F, int32_t H, int32_t W,
                         double* partials, float* out, void* stream);
/* Sharded form of the same gain: vl3d_frame_sum gives out[i] = sum_f v[f*chw + i] over a rank's block
 * of target frames; after an all-reduce(SUM) of those (3,H,W) sums, vl3d_scale_invariant_presum evaluates
 * MPV.py:499-504 with mean_F res = res_sum / F. */
int vl3d_frame_sum(const float* v, int32_t n_frames, int64_t chw, float* out, void* stream);
int vl3d_scale_invariant_presum(const float* rgb, int32_t T, const float* res_sum, int32_t F, int32_t H, int32_t W,
                                double* partials, float* out, void* stream);
/* Band-sharded form (the loss split over ranks by pixel-row bands): vl3d_scale_log_sum writes
 *   sum_out[0] = sum over channels, rows [row_begin,row_end) and columns of log((mean_F res + .01)/(mean_T rgb + .01))
 * for a band buffer rgb (>=T,3,H,W) / res (F,3,H,W); after an all-reduce(SUM) of the ranks' sums,
 * vl3d_scale_finish gives out[0] = (exp(log_sum[0] / count) + 3) / 4 with count = 3 * (pixels of the whole image). */
int vl3d_scale_log_sum(const float* rgb, int32_t T, const float* res, int32_t F, int32_t H, int32_t W,
                       int32_t row_begin, int32_t row_end, double* partials, double* sum_out, void* stream);
int vl3d_scale_finish(const double* log_sum, int64_t count, float* out, void* stream);
/* out[i] = x[i] * xscale[0] for n contiguous floats (rgb_pad * scale, MPV.py:504). */
int vl3d_scale_video(const float* x, const float* xscale, float* out, int64_t n, void* stream);

/* ---- looping loss (utils_vid.py) ---------------------------------------------------------------
 * vl3d_patchnn_search: extract_3Dpatches + efficient_compute_distances + get_col_mins_efficient +
 *   get_NN_indices_low_memory (utils_vid.py:60-142) fused; distances are direct sums of squared
 *   differences in fp32.  x must already carry the scale-invariant gain (vl3d_scale_video).
 *   nn_out: (ho, wo, n1) int32, first-minimum tie rule.  Only patch rows [row_begin, row_end) are
 *   searched and written (patch positions are independent: this is how ranks split the search).
 * vl3d_vote_loss: gather + FoldNd votes / counts + robust_lossfun + its derivative
 *   (utils_vid.py:217-229, 344-348, 10-26).  rou_kind: 0 general float rou, 1 'mse', 2 'abs'.
 *   y2x_out (3,t,h,w)/weight_out (t,h,w) optional (last_y2x / last_weight caches).
 *   grad_out: (Tx_full, 3, Hfull, Wfull) = gcoef * xscale * rho'(x*xscale - y2x) / N inside the
 *   fitted crop, 0 outside (optional).  Only frames [frame_begin, frame_end) and pixel rows [row_begin, row_end) of x
 *   contribute (ranks split the frames, or the rows: then x / y / nn describe the rank's row band and rows outside
 *   the range are the overlap owned by a neighbour; their gradient is written as 0); loss_out[0] = (sum of rho over
 *   those frames and rows) / N_total, N_total = n_total if > 0 else 3*t*h*w of `desc`, so partial results add up to the
 *   mean.  partials: workspace of >= vl3d_vote_partials(Tx_full, Hfull, Wfull) doubles. */
int vl3d_patchnn_search(const vl3d_loss_desc* desc, const float* x, const float* y,
                        int32_t row_begin, int32_t row_end, int32_t* nn_out, void* stream);
int vl3d_vote_partials(int32_t Tx_full, int32_t Hfull, int32_t Wfull);
int vl3d_vote_loss(const vl3d_loss_desc* desc, const float* x, const float* xscale, const float* y,
                   const int32_t* nn, int32_t rou_kind, float rou, float scaling, float gcoef,
                   int32_t Tx_full, int32_t Hfull, int32_t Wfull, int32_t frame_begin, int32_t frame_end,
                   int32_t row_begin, int32_t row_end, int64_t n_total,
                   float* y2x_out, float* weight_out, float* grad_out, double* partials, float* loss_out,
                   void* stream);

/* The two trivial entries of MPMeshVid.losses (MPV.py:135-136), value and dL/dx in one pass over channel-major videos
 * x (3,tx,h,w), y (3,ty,h,w), both contiguous:
 *   kind 0 = Patch3DMSE (utils_vid.py:437-440): mean over the first min(tx,ty) frames of (x - y)^2;
 *   kind 1 = Patch3DAvg (utils_vid.py:443-445): mean over (c,h,w) of (mean_t x - mean_t y)^2.
 * grad_out (3,tx,h,w) or NULL; partials: >= vl3d_video_loss_partials() doubles. */
int vl3d_video_loss_partials(void);
int vl3d_video_loss(int32_t kind, const float* x, const float* y, int32_t tx, int32_t ty, int32_t h, int32_t w,
                    float* grad_out, double* partials, float* loss_out, void* stream);

/* ---- "next" rows ------------------------------------------------------------
 * vl3d_patch_l1 (N3, evaluations/NNMSE.py:45-53): err_out[ho,wo,n1] = mean |y_patch[nn] - x_patch| over the
 *   3*pt*p*p elements of each (patch position, query) pair, given the NN map from vl3d_patchnn_search.
 * vl3d_to8b (N1, utils.py:17 `to8b`): rgb (T,3,H,W) float -> out (T,H,W,3) uint8 = trunc(255*clip(x,0,1)). */
int vl3d_patch_l1(const vl3d_loss_desc* desc, const float* x, const float* y, const int32_t* nn, float* err_out,
                  void* stream);
int vl3d_to8b(const float* rgb, uint8_t* out, int32_t T, int32_t H, int32_t W, void* stream);

/* S3 (train_3dvid.py:54 `vid / 255` of MVVidPatchDataset): uint8 frames -> fp32 in [0,1] on the device (IEEE division:
 * bit-identical to the host conversion), so that target videos cross PCIe / NVLink as bytes.
 * src: `planes` images (frames x channels) of H x W bytes, plane / row strides in bytes (a crop of a larger video);
 * dst: contiguous (planes, H, W) floats. */
int vl3d_u8_to_unit(const uint8_t* src, float* dst, int32_t planes, int32_t H, int32_t W, int64_t plane_stride,
                    int64_t row_stride, void* stream);

/* ---- optimiser (MPV.py:200-218: torch.optim.Adam(betas=(0.9,0.999), eps=6e-8), one tensor) ------
 * p, g, m, v: n floats each; step >= 1.  lr etc. are host scalars. */
int vl3d_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                   float beta1, float beta2, float eps, void* stream);

/* ---- fused backward + Adam (train_3dvid.py:242-244 `zero_grad / backward / step` for atlas_dyn) ----
 * One persistent kernel: the tiles of vl3d_composite_bwd (pad = 0, ts = NULL) and Adam (vl3d_adam_step's arithmetic)
 * on rectangles of atlas_dyn, pulled from ONE ordered work queue described by `items` (host-built, see
 * videoloop3d_b200/schedule.py; n_items x 12 int32 = {type | flags << 4, a, b, c; wait_first, wait_count, wait_target,
 * signal; wait2_first, wait2_count, wait2_target, 0}: type 0 = tile (a, b) of chunk frames, 1 = Adam / 2 = zero-gradient
 * on `c` rows x `b` texels starting at texel `a` of each frame of the chunk, row stride dyn_w; an item starts when
 * every counter of its two wait ranges has reached the range's target and bumps counter `signal` when done).  The table describes one round = one chunk of 2 frames and is
 * replayed for T/2 chunks (`n_rounds` = T/2, or T/2 + 1 when items are flagged "previous round").
 * T must be even.  counters: n_rounds' worth of n_counters int32 (initialised by the caller), ticket: one int32 (zeroed).
 * grad_dyn: in schedules whose Adam items re-zero it, it must be all-zero on entry and is all-zero on exit; in
 * zero-ahead schedules its content on entry / exit is irrelevant.  grad_sta is accumulated into as in vl3d_composite_bwd
 * (the static atlas is optimised by the caller after its all-reduce).  atlas_dyn, adam_m, adam_v are updated in place.
 * ctas_per_sm: 0 = as many as fit. */
int vl3d_fused_bwd_adam(const vl3d_view* view, const vl3d_quad* quads, float* atlas_dyn, const float* atlas_sta,
                        int32_t T, const float* grad_rgb, const float* rgb, const float* w_smooth,
                        double* smooth_sums, float* grad_dyn, float* grad_sta, float* adam_m, float* adam_v,
                        int32_t step, float lr, float beta1, float beta2, float eps, const int32_t* items,
                        int32_t n_items, int32_t n_rounds, int32_t* counters, int32_t n_counters,
                        int32_t* ticket, int32_t ctas_per_sm, void* stream);

/* ---- the same pass in "owner" mode (dense layout = VL3D_VIEW_RECT_PLANES, regulariser weights given) -------------------
 * A texel whose bilinear footprint is met by the pixels of one screen tile only gets its complete gradient inside that
 * tile: the tile accumulates it in a per-CTA scratch box (L2-resident), runs Adam on it there and never touches
 * grad_dyn; only texels near tile borders go through grad_dyn and the queue's ADAM items, which skip the owned texels.
 * Ownership is decided per (plane, texel) from `own`:
 *   hinv[d]  row-major 3x3: plane-local texel (x - rect[d].x0, y - rect[d].y0, 1) -> (px*w, py*w, w), the continuous pixel
 *            index of the texel centre in the target view (pixel p's centre = p), w > 0 where visible;
 *   reach[d] >= the largest distance (L-inf, pixels) between a pixel and the screen position of a texel it taps
 *            (>= 16 disables ownership on the plane);
 *   rect[d]  x0, y0, x1, y1 (inclusive): the atlas texels tapped by plane d and by no other plane.
 * ADAM items carry the plane of their rectangle in the 12th int32 of the item (-1 = no plane: nothing is skipped).
 * own_table: workspace of vl3d_fused_own_table_bytes(H, W) bytes (per screen tile the planes it owns and where; filled by
 * a small kernel in front of the pass); scratch: vl3d_fused_own_scratch_bytes(ctas) bytes for `ctas` resident CTAs
 * (SMs x 3 is always enough), ALL-ZERO on entry and all-zero on exit.  Both 16-byte aligned. */
typedef struct vl3d_own {
    float   hinv[VL3D_MAX_PLANES * 9];
    float   reach[VL3D_MAX_PLANES];
    int32_t rect[VL3D_MAX_PLANES * 4];
} vl3d_own;
int64_t vl3d_fused_own_scratch_bytes(int32_t ctas);
int64_t vl3d_fused_own_table_bytes(int32_t H, int32_t W);
int vl3d_fused_bwd_adam_own(const vl3d_view* view, const vl3d_quad* quads, float* atlas_dyn, const float* atlas_sta,
                            int32_t T, const float* grad_rgb, const float* rgb, const float* w_smooth,
                            double* smooth_sums, float* grad_dyn, float* grad_sta, float* adam_m, float* adam_v,
                            int32_t step, float lr, float beta1, float beta2, float eps, const int32_t* items,
                            int32_t n_items, int32_t n_rounds, int32_t* counters, int32_t n_counters,
                            int32_t* ticket, int32_t ctas_per_sm, const vl3d_own* own, uint32_t* own_table,
                            int64_t own_table_bytes, float* scratch, int64_t scratch_bytes, void* stream);

/* ---- exchanges of the T-sharded step over peer memory ----------------------------------------------
 * vl3d_copy_boxes: up to VL3D_MAX_BOXES strided box copies in ONE launch.  A box is (n_frames, n_planes, n_rows, n_cols)
 * floats with element strides (frame, plane, row) on each side, columns contiguous; dst may be memory of a peer GPU mapped
 * into this process (NVLink stores).  src2 (optional, same strides as src) is added on the fly (adjoint of the loop pad,
 * MPV.py:490-492); without it the copy is bit-exact for any 32-bit payload.  `boxes` is HOST memory (copied into the launch). */
#define VL3D_MAX_BOXES 32
typedef struct vl3d_box {
    const float* src;
    const float* src2;
    float*       dst;
    int32_t n_frames, n_planes, n_rows, n_cols;
    int64_t src_sf, src_sp, src_sr;
    int64_t dst_sf, dst_sp, dst_sr;
} vl3d_box;
int vl3d_copy_boxes(const vl3d_box* boxes, int32_t n_boxes, void* stream);

/* ---- optional placement helper ------------------------------------------------------------------------------------------
 * Device memory with the GENERIC compression attribute (CUDA virtual memory management), for buffers that are mostly
 * zeros when they cross HBM (grad_dyn of vl3d_fused_bwd_adam: written back as zeros by the Adam items, fetched again by
 * the first RED).  Results never depend on it.  Fails with VL3D_EINVAL when the device / driver does not support it.
 * bytes_out: the mapped size (rounded up to the allocation granularity); compressed_out: 1 if the driver granted the
 * attribute.  The only state the library keeps besides the last error: the handles needed by vl3d_free_compressible. */
int vl3d_alloc_compressible(int64_t bytes, void** ptr_out, int64_t* bytes_out, int32_t* compressed_out);
int vl3d_free_compressible(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* VL3D_H_ */
