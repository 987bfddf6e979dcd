"""Thin torch-facing wrappers over the C ABI (include/vl3d.h) + the autograd Functions built on them.

PyTorch is only plumbing here: device memory, the current stream and autograd bookkeeping.  Every
numerical operation of the hot path runs inside libvl3d.so; nothing falls back to eager PyTorch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from . import tiles


def _require_cuda(t, name):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.Vl3dError(f"{name} must be a CUDA tensor: the vl3d hot path has no CPU implementation")
    if t.dtype != torch.float32 and t.dtype not in (torch.int32, torch.float64):
        raise _lib.Vl3dError(f"{name}: unsupported dtype {t.dtype} (the reference's fp16 flag is 'do NOT use')")


def as_texels(atlas):
    """(B,4,H,W) logical tensor whose memory is RGBA-interleaved (channels_last).  Returns the tensor
    itself when it already is, else a converted copy (the caller should then re-home the parameter)."""
    if atlas.dim() != 4 or atlas.shape[1] != 4:
        raise _lib.Vl3dError(f"atlas must be (B,4,H,W), got {tuple(atlas.shape)} (atlas_cnl must be 4, rgb_mlp_type=direct)")
    b, c, h, w = atlas.shape
    want = (h * w * 4, 1, w * 4, 4)
    if tuple(atlas.stride()) == want:
        return atlas
    out = torch.empty_strided((b, c, h, w), want, dtype=atlas.dtype, device=atlas.device)
    out.copy_(atlas)
    return out


# ------------------------------------------------------------------------------------------------
# view pack
# ------------------------------------------------------------------------------------------------
@dataclass
class MeshPack:
    """Device-resident quad table + per-plane vertex grids (rebuilt on lod()/load only)."""
    quads: torch.Tensor          # uint8 view of the vl3d_quad array, on device
    grids: np.ndarray            # (D,5) X0,Y0,dX,dY,z (float64, host)
    D: int
    qh: int
    qw: int
    n_static: int
    n_dynamic: int
    rect_planes: bool = False    # dense layout: one atlas rectangle per plane (enables the TMA render)
    table: np.ndarray = None     # host copy of the quad table (schedule.band_schedule reads the plane rectangles)


def make_mesh_pack(model_tensors, D, hv, wv, device):
    t = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in model_tensors.items()}
    grids = tiles.plane_grids(t["verts"], D, hv, wv)
    table = tiles.build_quad_table(D, hv, wv, t["faces"], t["uvs"], t["uvfaces"], t["atlas_hw"],
                                   t["faces_dyn"], t["uvs_dyn"], t["uvfaces_dyn"], t["atlas_dyn_hw"])
    q = torch.from_numpy(table.view(np.uint8).copy()).to(device)
    return MeshPack(quads=q, grids=grids, D=D, qh=hv - 1, qw=wv - 1,
                    n_static=int((table["kind"] == 1).sum()), n_dynamic=int((table["kind"] == 2).sum()),
                    rect_planes=tiles.planes_are_rectangles(table, D, hv - 1, wv - 1), table=table)


def make_view(pack: MeshPack, H, W, tar_extrin, tar_intrin, ref_extrin, dyn_hw, sta_hw):
    """Host-side vl3d_view for one target camera (float64 maths, see tiles.view_homographies)."""
    to_np = lambda a: a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    homs, cx, cy = tiles.view_homographies(pack.grids, pack.qh, pack.qw, to_np(tar_extrin), to_np(tar_intrin),
                                           to_np(ref_extrin), H, W)
    v = _lib.View()
    v.H, v.W, v.D, v.qh, v.qw = int(H), int(W), pack.D, pack.qh, pack.qw
    v.dyn_h, v.dyn_w = int(dyn_hw[0]), int(dyn_hw[1])
    v.sta_h, v.sta_w = int(sta_hw[0]), int(sta_hw[1])
    v.cx, v.cy = cx, cy
    v.flags = _lib.VIEW_RECT_PLANES if pack.rect_planes else 0
    flat = homs.reshape(-1)
    C.memmove(v.hom, flat.ctypes.data, flat.nbytes)
    return v


# ------------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------------
def composite_fwd(view, pack, atlas_dyn, atlas_sta, ts, T, pad, rgb_out=None, want_alpha=False, smooth_sums=None,
                  want_mpi=False, want_hits=False):
    dev = atlas_dyn.device
    _require_cuda(atlas_dyn, "atlas_dyn"); _require_cuda(atlas_sta, "atlas")
    H, W = view.H, view.W
    if rgb_out is None:
        rgb_out = torch.empty((T + pad, 3, H, W), dtype=torch.float32, device=dev)
    alpha = torch.empty((T, H, W), dtype=torch.float32, device=dev) if want_alpha else None
    mpi = torch.zeros((T, H, W, view.D, 4), dtype=torch.float32, device=dev) if want_mpi else None
    hits = torch.zeros((H, W), dtype=torch.int32, device=dev) if want_hits else None
    _lib.call("vl3d_composite_fwd", C.byref(view), _lib.ptr(pack.quads), _lib.ptr(atlas_dyn), _lib.ptr(atlas_sta),
              _lib.ptr(ts), int(T), int(pad), _lib.ptr(rgb_out), _lib.ptr(alpha), _lib.ptr(smooth_sums),
              _lib.ptr(mpi), _lib.ptr(hits), _lib.stream_ptr())
    return rgb_out, alpha, mpi, hits


def composite_bwd(view, pack, atlas_dyn, atlas_sta, ts, T, pad, grad_rgb, rgb, w_smooth, grad_dyn, grad_sta,
                  smooth_sums=None):
    _lib.call("vl3d_composite_bwd", C.byref(view), _lib.ptr(pack.quads), _lib.ptr(atlas_dyn), _lib.ptr(atlas_sta),
              _lib.ptr(ts), int(T), int(pad), _lib.ptr(grad_rgb), _lib.ptr(rgb), _lib.ptr(w_smooth),
              _lib.ptr(smooth_sums), _lib.ptr(grad_dyn), _lib.ptr(grad_sta), _lib.stream_ptr())


def make_inv_depth(pack: MeshPack, H, W, tar_extrin, tar_intrin, ref_extrin, scale=1.0, offset=0.0):
    """Host (D,3) float32 coefficients of `scale / view depth + offset` per plane (tiles.view_inv_depth)."""
    to_np = lambda a: a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(tiles.view_inv_depth(pack.grids, to_np(tar_extrin), to_np(tar_intrin), to_np(ref_extrin), H, W,
                                                     scale, offset))


def _inv_depth_arg(view, a):
    """(array kept alive by the caller, pointer) for the `inv_depth_host` argument: 3*D contiguous float32 on the host."""
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size < 3 * view.D:
        raise _lib.Vl3dError(f"inv_depth needs 3 coefficients per plane ({3 * view.D} floats), got {a.size}")
    return a, a.ctypes.data_as(C.c_void_p)


def composite_terms_fwd(view, pack, atlas_dyn, atlas_sta, ts, T, inv_depth=None, sparsity_eps=1e-4, want_alpha=True,
                        want_disp=False, want_sparsity=False):
    """alpha (T,H,W), disp (T,H,W), sparsity sum (1,) float64 of vl3d_composite_terms_fwd (each None unless wanted)."""
    _require_cuda(atlas_dyn, "atlas_dyn"); _require_cuda(atlas_sta, "atlas")
    dev = atlas_dyn.device
    H, W = view.H, view.W
    if want_disp and inv_depth is None:
        raise _lib.Vl3dError("composite_terms_fwd: disp needs the inverse-depth coefficients")
    alpha = torch.empty((T, H, W), dtype=torch.float32, device=dev) if want_alpha else None
    disp = torch.empty((T, H, W), dtype=torch.float32, device=dev) if want_disp else None
    sp = torch.zeros(1, dtype=torch.float64, device=dev) if want_sparsity else None
    keep, inv_ptr = _inv_depth_arg(view, inv_depth)
    _lib.call("vl3d_composite_terms_fwd", C.byref(view), _lib.ptr(pack.quads), _lib.ptr(atlas_dyn), _lib.ptr(atlas_sta),
              _lib.ptr(ts), int(T), inv_ptr, float(sparsity_eps), _lib.ptr(alpha), _lib.ptr(disp), _lib.ptr(sp),
              _lib.stream_ptr())
    del keep
    return alpha, disp, sp


def composite_terms_bwd(view, pack, atlas_dyn, atlas_sta, ts, T, inv_depth, sparsity_eps, g_alpha, g_disp, w_sparsity,
                        grad_dyn, grad_sta):
    for name, g in (("g_alpha", g_alpha), ("g_disp", g_disp)):
        if g is not None and (tuple(g.shape) != (T, view.H, view.W) or not g.is_contiguous() or g.dtype != torch.float32):
            raise _lib.Vl3dError(f"composite_terms_bwd: {name} must be a contiguous float32 (T,H,W) tensor, got {tuple(g.shape)} {g.dtype}")
    keep, inv_ptr = _inv_depth_arg(view, inv_depth)
    _lib.call("vl3d_composite_terms_bwd", C.byref(view), _lib.ptr(pack.quads), _lib.ptr(atlas_dyn), _lib.ptr(atlas_sta),
              _lib.ptr(ts), int(T), inv_ptr, float(sparsity_eps), _lib.ptr(g_alpha), _lib.ptr(g_disp),
              _lib.ptr(w_sparsity), _lib.ptr(grad_dyn), _lib.ptr(grad_sta), _lib.stream_ptr())
    del keep


def scale_invariant(rgb, T, res, out=None, partials=None):
    """MPV.py:499-504 on device: rgb (>=T,3,H,W) contiguous, res (F,3,H,W) contiguous -> (1,) float."""
    _require_cuda(rgb, "rgb"); _require_cuda(res, "res")
    F_, _, H, W = res.shape
    if not res.is_contiguous():
        res = res.contiguous()
    assert rgb.is_contiguous() and tuple(rgb.shape[1:]) == (3, H, W)
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=rgb.device)
    if partials is None:
        partials = torch.empty(_lib.load().vl3d_scale_partials(), dtype=torch.float64, device=rgb.device)
    _lib.call("vl3d_scale_invariant", _lib.ptr(rgb), int(T), _lib.ptr(res), int(F_), int(H), int(W),
              _lib.ptr(partials), _lib.ptr(out), _lib.stream_ptr())
    return out


def frame_sum(v, out=None):
    """sum over the leading (frame) axis of a contiguous (n,3,H,W) block -> (3,H,W)."""
    _require_cuda(v, "v")
    assert v.is_contiguous()
    if out is None:
        out = torch.empty(v.shape[1:], dtype=torch.float32, device=v.device)
    _lib.call("vl3d_frame_sum", _lib.ptr(v), int(v.shape[0]), int(v[0].numel()) if v.shape[0] else int(out.numel()),
              _lib.ptr(out), _lib.stream_ptr())
    return out


def scale_invariant_presum(rgb, T, res_sum, F_, out=None, partials=None):
    """MPV.py:499-504 from the frame-summed target `res_sum` (3,H,W) (the sharded step all-reduces it)."""
    _require_cuda(rgb, "rgb"); _require_cuda(res_sum, "res_sum")
    _, H, W = res_sum.shape
    assert rgb.is_contiguous() and res_sum.is_contiguous() and tuple(rgb.shape[1:]) == (3, H, W)
    if out is None:
        out = torch.empty(1, dtype=torch.float32, device=rgb.device)
    if partials is None:
        partials = torch.empty(_lib.load().vl3d_scale_partials(), dtype=torch.float64, device=rgb.device)
    _lib.call("vl3d_scale_invariant_presum", _lib.ptr(rgb), int(T), _lib.ptr(res_sum), int(F_), int(H), int(W),
              _lib.ptr(partials), _lib.ptr(out), _lib.stream_ptr())
    return out


def scale_log_sum(rgb, T, res, rows, partials, out):
    """Band-sharded scale-invariant gain, step 1 (MPV.py:499-504): out[0] (float64) = sum over channels, `rows` and columns
    of log((mean_F res + .01)/(mean_T rgb + .01)); rgb (>=T,3,H,W) and res (F,3,H,W) contiguous band buffers."""
    _require_cuda(rgb, "rgb"); _require_cuda(res, "res")
    F_, _, H, W = res.shape
    assert rgb.is_contiguous() and res.is_contiguous() and tuple(rgb.shape[1:]) == (3, H, W)
    _lib.call("vl3d_scale_log_sum", _lib.ptr(rgb), int(T), _lib.ptr(res), int(F_), int(H), int(W), int(rows[0]), int(rows[1]),
              _lib.ptr(partials), _lib.ptr(out), _lib.stream_ptr())
    return out


def scale_finish(log_sum, count, out):
    """Step 2: out[0] = (exp(log_sum[0] / count) + 3) / 4."""
    _lib.call("vl3d_scale_finish", _lib.ptr(log_sum), int(count), _lib.ptr(out), _lib.stream_ptr())
    return out


def _fit(size, p, st, name):
    """fit_patch of utils_vid.py:307-313."""
    if size < p:
        raise ValueError(f"{name}={size} is smaller than the patch size {p}")
    return (size - p) // st * st + p if (size - p) % st != 0 else size


def make_loss_desc(x_tchw_shape, x_strides, y_fchw_shape, y_strides, patch_size, patcht_size, stride, stridet, alpha,
                   fit=True):
    """vl3d_loss_desc for x (t,3,h,w)-indexed and y (F,3,h,w)-indexed planar videos.
    strides: (frame, channel, row) element strides; the pixel stride must be 1."""
    p, pt, s, st = int(patch_size), int(patcht_size), int(stride), int(stridet)
    tx, _, h, w = x_tchw_shape
    F_ = y_fchw_shape[0]
    d = _lib.LossDesc()
    if fit:
        d.t, d.h, d.w = _fit(tx, pt, st, "frame_num"), _fit(h, p, s, "patch_height"), _fit(w, p, s, "patch_width")
    else:                                            # Patch3DGPNNDirectLoss: unfold's floor semantics
        if tx < pt or h < p or w < p:
            raise ValueError("video smaller than the patch")
        d.t, d.h, d.w = int(tx), int(h), int(w)
    d.F = int(F_)
    if F_ < pt:
        raise ValueError(f"target video has {F_} frames < patcht_size {pt}")
    d.p, d.pt, d.s, d.st = p, pt, s, st
    d.n1, d.n2 = (d.t - pt) // st + 1, (F_ - pt) // st + 1
    d.ho, d.wo = (d.h - p) // s + 1, (d.w - p) // s + 1
    d.x_sf, d.x_sc, d.x_sr = (int(v) for v in x_strides)
    d.y_sf, d.y_sc, d.y_sr = (int(v) for v in y_strides)
    alpha = float(alpha)
    d.use_alpha = 0 if alpha > 100 else 1            # utils_vid.py:208
    d.alpha = 0.0 if alpha > 100 else alpha
    return d


def u8_to_unit(v, out=None):
    """uint8 frames (F,3,h,w) — contiguous or a spatial crop of a contiguous video — -> float32 `v / 255`
    (train_3dvid.py:54), bit-identical to the host conversion."""
    if not v.is_cuda:
        raise _lib.Vl3dError("v must be a CUDA tensor: the vl3d hot path has no CPU implementation")
    if v.dtype != torch.uint8 or v.dim() != 4:
        raise _lib.Vl3dError(f"u8_to_unit expects a (F,3,h,w) uint8 tensor, got {tuple(v.shape)} {v.dtype}")
    F_, C_, h, w = v.shape
    if v.stride(3) != 1 or (F_ > 1 and v.stride(0) != C_ * v.stride(1)):
        v = v.contiguous()
    if out is None or tuple(out.shape) != (F_, C_, h, w) or out.dtype != torch.float32 or not out.is_contiguous():
        out = torch.empty((F_, C_, h, w), dtype=torch.float32, device=v.device)
    _lib.call("vl3d_u8_to_unit", _lib.ptr(v), _lib.ptr(out), int(F_ * C_), int(h), int(w), int(v.stride(1)),
              int(v.stride(2)), _lib.stream_ptr())
    return out


def scale_video(x, xscale, out=None):
    """x * xscale into a contiguous buffer (MPV.py:504)."""
    if not x.is_contiguous():
        x = x.contiguous()
    if out is None or out.shape != x.shape:
        out = torch.empty_like(x)
    _lib.call("vl3d_scale_video", _lib.ptr(x), _lib.ptr(xscale), _lib.ptr(out), int(x.numel()), _lib.stream_ptr())
    return out


def patchnn_search(desc, x, xscale, y, nn_out=None, rows=None, scaled_ws=None):
    """NN indices (ho,wo,n1) int32.  With a scale-invariant gain the search runs on a pre-scaled copy of x
    (`scaled_ws`: optional persistent workspace of x's shape)."""
    if nn_out is None:
        nn_out = torch.empty((desc.ho, desc.wo, desc.n1), dtype=torch.int32, device=x.device)
    r0, r1 = (0, desc.ho) if rows is None else rows
    if xscale is not None:
        x = scale_video(x, xscale, out=scaled_ws)
        d2 = _lib.LossDesc.from_buffer_copy(desc)
        d2.x_sf, d2.x_sc, d2.x_sr = x.stride(0), x.stride(1), x.stride(2)
        desc = d2
    _lib.call("vl3d_patchnn_search", C.byref(desc), _lib.ptr(x), _lib.ptr(y), int(r0), int(r1),
              _lib.ptr(nn_out), _lib.stream_ptr())
    return nn_out


def parse_rou(rou):
    """robust_lossfun's rou (utils_vid.py:10-16) arrives as str from the config."""
    if rou == "mse":
        return 1, 0.0
    if rou == "abs":
        return 2, 0.0
    return 0, float(rou)


def vote_loss(desc, x, xscale, y, nn, rou, scaling, gcoef, full_shape, want_cache=False, grad_out=None,
              want_grad=True, partials=None, loss_out=None, frames=None, rows=None, n_total=0):
    """`frames` / `rows`: the frame range / pixel-row range this call owns (default: everything); `n_total`: element
    count of the whole problem when `desc` describes a rank's row band (0: 3*t*h*w of `desc`)."""
    Tx, Hf, Wf = full_shape
    dev = x.device
    kind, rouf = parse_rou(rou)
    y2x = torch.empty((1, 3, desc.t, desc.h, desc.w), dtype=torch.float32, device=dev) if want_cache else None
    wgt = torch.empty((1, 1, desc.t, desc.h, desc.w), dtype=torch.float32, device=dev) if want_cache else None
    if want_grad and grad_out is None:
        grad_out = torch.empty((Tx, 3, Hf, Wf), dtype=torch.float32, device=dev)
    n_part = _lib.load().vl3d_vote_partials(int(Tx), int(Hf), int(Wf))
    if partials is None or partials.numel() < n_part:
        partials = torch.empty(n_part, dtype=torch.float64, device=dev)
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    _lib.call("vl3d_vote_loss", C.byref(desc), _lib.ptr(x), _lib.ptr(xscale), _lib.ptr(y), _lib.ptr(nn),
              int(kind), float(rouf), float(scaling), float(gcoef), int(Tx), int(Hf), int(Wf),
              int(0 if frames is None else frames[0]), int(Tx if frames is None else frames[1]),
              int(0 if rows is None else rows[0]), int(Hf if rows is None else rows[1]), int(n_total),
              _lib.ptr(y2x), _lib.ptr(wgt), _lib.ptr(grad_out if want_grad else None), _lib.ptr(partials),
              _lib.ptr(loss_out), _lib.stream_ptr())
    return loss_out, grad_out, y2x, wgt


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=6e-8):
    """In-place Adam on tensors that share one dense memory layout (MPV.py:213)."""
    for t in (g, m, v):
        if tuple(t.stride()) != tuple(p.stride()) or t.shape != p.shape:
            raise _lib.Vl3dError("adam_step: p, g, m, v must share shape and strides")
    _lib.call("vl3d_adam_step", _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), int(p.numel()), int(step),
              float(lr), float(beta1), float(beta2), float(eps), _lib.stream_ptr())


def fused_bwd_adam(view, pack, atlas_dyn, atlas_sta, T, grad_rgb, rgb, w_smooth, smooth_sums, grad_dyn, grad_sta, m, v,
                   step, lr, beta1, beta2, eps, items, n_items, n_rounds, state, n_counters, ctas_per_sm=0):
    """Backward + Adam of `atlas_dyn[:T]` in one persistent kernel (csrc/fused_bwd_adam.cu).  `items`: device int32
    (n_items, 12) table from schedule.py; `state`: device int32 scratch prepared by the caller: [0] = queue head (0),
    [16:] = n_rounds x n_counters counters (each round initialised to the schedule's `counter_init`)."""
    for t in (grad_dyn, m, v):
        if tuple(t.stride()) != tuple(atlas_dyn.stride()) or t.shape != atlas_dyn.shape:
            raise _lib.Vl3dError("fused_bwd_adam: atlas_dyn, grad_dyn, m, v must share shape and strides")
    if (items.dtype != torch.int32 or not items.is_contiguous() or items.numel() != 12 * n_items or state.dtype != torch.int32
            or state.numel() < 16 + n_rounds * n_counters):
        raise _lib.Vl3dError("fused_bwd_adam: bad schedule buffers")
    _lib.call("vl3d_fused_bwd_adam", C.byref(view), _lib.ptr(pack.quads), _lib.ptr(atlas_dyn), _lib.ptr(atlas_sta), int(T),
              _lib.ptr(grad_rgb), _lib.ptr(rgb), _lib.ptr(w_smooth), _lib.ptr(smooth_sums), _lib.ptr(grad_dyn),
              _lib.ptr(grad_sta), _lib.ptr(m), _lib.ptr(v), int(step), float(lr), float(beta1), float(beta2), float(eps),
              _lib.ptr(items), int(n_items), int(n_rounds), C.c_void_p(state.data_ptr() + 64), int(n_counters),
              _lib.ptr(state), int(ctas_per_sm), _lib.stream_ptr())


class _CompressibleBlock:
    """A vl3d_alloc_compressible allocation exposed through __cuda_array_interface__ (freed with the last tensor view)."""

    def __init__(self, n_floats, device):
        ptr, size, comp = C.c_void_p(), C.c_int64(), C.c_int32()
        with torch.cuda.device(device):
            _lib.call("vl3d_alloc_compressible", int(n_floats) * 4, C.byref(ptr), C.byref(size), C.byref(comp))
        self.ptr, self.bytes, self.compressed, self.device = ptr.value, size.value, bool(comp.value), device
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (self.ptr, False), "version": 3,
                                         "strides": None}

    def __del__(self):
        try:
            _lib.load().vl3d_free_compressible(C.c_void_p(self.ptr))
        except Exception:
            pass


def compressible_zeros_like(ref):
    """All-zero float32 tensor with `ref`'s shape and strides in compressible device memory (include/vl3d.h: placement
    helper), or None when the device / driver does not offer it.  Returns (tensor, granted)."""
    try:
        n = max(1 + sum((s - 1) * st for s, st in zip(ref.shape, ref.stride())), 1) if ref.numel() else 1
        blk = _CompressibleBlock(n, ref.device)
    except _lib.Vl3dError:
        return None, False
    with torch.cuda.device(ref.device):
        flat = torch.as_tensor(blk, device=ref.device)
        if flat.data_ptr() != blk.ptr:                              # (torch copied instead of wrapping: no point)
            return None, False
        flat.zero_()
        t = flat.as_strided(tuple(ref.shape), tuple(ref.stride()))
    t._vl3d_block = blk                                            # keep the allocation alive with the view
    return t, blk.compressed


def copy_boxes(boxes):
    """One launch of strided 4-D box copies (csrc/exchange.cu).  `boxes`: list of (src, dst, src2) with src / src2 4-D
    float32 (or int32) tensors whose last dimension is contiguous and dst a tensor of the same shape — possibly a view
    of a PEER GPU's symmetric-memory buffer (then the copy is a stream of NVLink stores)."""
    if not boxes:
        return
    for k in range(0, len(boxes), _lib.MAX_BOXES):
        part = boxes[k:k + _lib.MAX_BOXES]
        arr = (_lib.Box * len(part))()
        for b, (src, dst, src2) in zip(arr, part):
            if src.dim() != 4 or tuple(dst.shape) != tuple(src.shape) or src.element_size() != 4 or dst.element_size() != 4:
                raise _lib.Vl3dError("copy_boxes: src / dst must be 4-D tensors of 32-bit elements with equal shapes")
            if (src.numel() and (src.stride(3) != 1 or dst.stride(3) != 1)) or (src2 is not None and src2.stride() != src.stride()):
                raise _lib.Vl3dError("copy_boxes: columns must be contiguous and src2 strided like src")
            b.src, b.dst = src.data_ptr(), dst.data_ptr()
            b.src2 = src2.data_ptr() if src2 is not None else None
            b.n_frames, b.n_planes, b.n_rows, b.n_cols = (int(v) for v in src.shape)
            b.src_sf, b.src_sp, b.src_sr = (int(v) for v in src.stride()[:3])
            b.dst_sf, b.dst_sp, b.dst_sr = (int(v) for v in dst.stride()[:3])
        _lib.call("vl3d_copy_boxes", arr, len(part), _lib.stream_ptr())


def fused_own_scratch(device, ctas=None):
    """All-zero scratch for `vl3d_fused_bwd_adam_own` (one box pair per resident CTA; the kernel leaves it all-zero)."""
    if ctas is None:
        ctas = 3 * torch.cuda.get_device_properties(device).multi_processor_count
    n = int(_lib.load().vl3d_fused_own_scratch_bytes(int(ctas)))
    return torch.zeros(n // 4, dtype=torch.float32, device=device)


def fused_own_table(device, H, W):
    """Workspace for the ownership tables of `vl3d_fused_bwd_adam_own` (filled by the call itself)."""
    return torch.empty(int(_lib.load().vl3d_fused_own_table_bytes(int(H), int(W))) // 4, dtype=torch.int32, device=device)


def fused_bwd_adam_own(view, pack, atlas_dyn, atlas_sta, T, grad_rgb, rgb, w_smooth, smooth_sums, grad_dyn, grad_sta, m, v,
                       step, lr, beta1, beta2, eps, items, n_items, n_rounds, state, n_counters, own, own_table, scratch,
                       ctas_per_sm=0):
    """`fused_bwd_adam` in owner mode (dense layout): `own` = `_lib.Own` from `tiles.own_descriptor`, `own_table` = int32
    workspace with one entry per screen tile, `scratch` from `fused_own_scratch` (ADAM items carry their plane)."""
    for t in (grad_dyn, m, v):
        if tuple(t.stride()) != tuple(atlas_dyn.stride()) or t.shape != atlas_dyn.shape:
            raise _lib.Vl3dError("fused_bwd_adam_own: atlas_dyn, grad_dyn, m, v must share shape and strides")
    if (items.dtype != torch.int32 or not items.is_contiguous() or items.numel() != 12 * n_items or state.dtype != torch.int32
            or state.numel() < 16 + n_rounds * n_counters):
        raise _lib.Vl3dError("fused_bwd_adam_own: bad schedule buffers")
    if own_table.dtype != torch.int32 or scratch.dtype != torch.float32:
        raise _lib.Vl3dError("fused_bwd_adam_own: bad own_table / scratch")
    _lib.call("vl3d_fused_bwd_adam_own", C.byref(view), _lib.ptr(pack.quads), _lib.ptr(atlas_dyn), _lib.ptr(atlas_sta), int(T),
              _lib.ptr(grad_rgb), _lib.ptr(rgb), _lib.ptr(w_smooth), _lib.ptr(smooth_sums), _lib.ptr(grad_dyn),
              _lib.ptr(grad_sta), _lib.ptr(m), _lib.ptr(v), int(step), float(lr), float(beta1), float(beta2), float(eps),
              _lib.ptr(items), int(n_items), int(n_rounds), C.c_void_p(state.data_ptr() + 64), int(n_counters),
              _lib.ptr(state), int(ctas_per_sm), C.byref(own), _lib.ptr(own_table), int(own_table.numel() * 4), _lib.ptr(scratch),
              int(scratch.numel() * 4), _lib.stream_ptr())


# ------------------------------------------------------------------------------------------------
# autograd Functions (the reference-compatible path: loss.backward(); optimizer.step())
# ------------------------------------------------------------------------------------------------
class CompositeFn(torch.autograd.Function):
    """rgb_pad (T+pad,3,H,W), alpha (T,H,W), smooth_sums (4,) float64 = f(atlas_dyn, atlas)."""

    @staticmethod
    def forward(ctx, atlas_dyn, atlas_sta, view, pack, ts, T, pad, smooth):
        ctx.set_materialize_grads(False)
        sums = torch.zeros(4, dtype=torch.float64, device=atlas_dyn.device) if smooth else None
        rgb, alpha, _, _ = composite_fwd(view, pack, atlas_dyn, atlas_sta, ts, T, pad, want_alpha=True,
                                         smooth_sums=sums)
        ctx.save_for_backward(atlas_dyn, atlas_sta, rgb, ts)
        ctx.view, ctx.pack, ctx.T, ctx.pad, ctx.smooth = view, pack, T, pad, smooth
        if sums is None:
            sums = torch.zeros(4, dtype=torch.float64, device=atlas_dyn.device)
        ctx.mark_non_differentiable(alpha)
        return rgb, alpha, sums

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_sums):
        atlas_dyn, atlas_sta, rgb, ts = ctx.saved_tensors
        if g_rgb is None:
            g_rgb = torch.zeros_like(rgb)
        if not g_rgb.is_contiguous():
            g_rgb = g_rgb.contiguous()
        w = None
        if ctx.smooth and g_sums is not None:
            w = g_sums.to(torch.float32).contiguous()
        grad_dyn = torch.zeros_like(atlas_dyn)     # preserves the channels_last texel layout
        grad_sta = torch.zeros_like(atlas_sta)
        composite_bwd(ctx.view, ctx.pack, atlas_dyn, atlas_sta, ts, ctx.T, ctx.pad, g_rgb, rgb, w, grad_dyn, grad_sta)
        return grad_dyn, grad_sta, None, None, None, None, None, None


class CompositeTermsFn(torch.autograd.Function):
    """alpha (T,H,W), disp (T,H,W), sparsity_sum (1,) float64 = f(atlas_dyn, atlas): the optional, differentiable per-ray
    terms of the composite (MPV.py:454-466, 511-515; csrc/terms.cu).  Outputs that are not wanted come back as zeros."""

    @staticmethod
    def forward(ctx, atlas_dyn, atlas_sta, view, pack, ts, T, inv_depth, sparsity_eps, want_disp, want_sparsity):
        ctx.set_materialize_grads(False)
        alpha, disp, sp = composite_terms_fwd(view, pack, atlas_dyn, atlas_sta, ts, T, inv_depth, sparsity_eps,
                                              want_alpha=True, want_disp=want_disp, want_sparsity=want_sparsity)
        ctx.save_for_backward(atlas_dyn, atlas_sta, ts)
        ctx.view, ctx.pack, ctx.T, ctx.inv_depth, ctx.eps = view, pack, T, inv_depth, sparsity_eps
        ctx.want_disp, ctx.want_sparsity = want_disp, want_sparsity
        if disp is None:
            disp = torch.zeros((), dtype=torch.float32, device=atlas_dyn.device)
        if sp is None:
            sp = torch.zeros(1, dtype=torch.float64, device=atlas_dyn.device)
        return alpha, disp, sp

    @staticmethod
    def backward(ctx, g_alpha, g_disp, g_sp):
        atlas_dyn, atlas_sta, ts = ctx.saved_tensors
        g_alpha = None if g_alpha is None else g_alpha.to(torch.float32).contiguous()
        g_disp = None if (g_disp is None or not ctx.want_disp) else g_disp.to(torch.float32).contiguous()
        w = None if (g_sp is None or not ctx.want_sparsity) else g_sp.to(torch.float32).contiguous()
        grad_dyn = torch.zeros_like(atlas_dyn)     # preserves the channels_last texel layout
        grad_sta = torch.zeros_like(atlas_sta)
        composite_terms_bwd(ctx.view, ctx.pack, atlas_dyn, atlas_sta, ts, ctx.T, ctx.inv_depth, ctx.eps, g_alpha, g_disp, w,
                            grad_dyn, grad_sta)
        return grad_dyn, grad_sta, None, None, None, None, None, None, None, None
