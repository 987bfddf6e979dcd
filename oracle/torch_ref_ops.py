"""ORACLE-SIDE BASELINE (test / measurement infrastructure, NOT product code) — the reference's own
*torch operator sequence* for one stage-2 step, device-agnostic, so that `bench.py --impl reference
--ref-device cuda` can time "the reference's single-GPU PyTorch render+loss step" (BASELINE.json north_star)
on the same B200 next to the CUDA path.  Only `tests/` and `bench.py --impl reference` import this module.

`oracle/mpv_oracle.py` restates the *mathematics* (its own bilinear filter, index_put, fp64); this module
instead issues the operators the reference issues, in the reference's order, so the time it takes is the
reference's time:

  render     MPV.py:406-454   F.grid_sample(atlas, uv.expand(T,1,N,2), zeros, align_corners=True) for the static
                              and the dynamic atlas (`atlas_dyn[ts]`: the index copy of the whole atlas, :439),
                              sigmoid, zeros canvas + two torch.masked_scatter, utils_mpi.py:92-107 overcompose
                              (cumprod / cat / mul / sum)
  forward    MPV.py:484-531   permute, loop pad by cat, scale-invariant gain, loss call, slot-wise rgb / alpha
                              smoothness on the (T,H,W,K,4) `mpi` tensor
  loss       utils_vid.py:294-349 macro-block loop; :206-229 unfold -> permute/reshape -> NN -> gather -> fold;
                              :72-86 expanded-form distances with a batched matmul; :109-142 column minimum and
                              argmin in chunks of 1024
  step       train_3dvid.py:230-244  weighted sum, zero_grad / backward / torch.optim.Adam(eps=6e-8) step

Two pieces cannot be the reference's: pytorch3d's rasteriser (absent; `mpv_oracle.geometry` supplies
`pix_to_face` / barycentric uv on the host, no_grad, not timed — the reference spends extra time there) and
unfoldNd (absent; UnfoldNd is restated as Tensor.unfold + reshape, FoldNd as index_add_ over the unfolded
index map, both cheaper than unfoldNd's one-hot convolution).  Both substitutions favour the baseline.

Checked against `mpv_oracle.forward_train` in tests/test_oracle.py (losses 1e-4 rel in fp32, NN map identical
up to fp32 near-ties).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import looploss_oracle as LL
from . import mpv_oracle as MO


# ------------------------------------------------------------------------------------------------
# host side, no_grad: what the rasteriser + get_uvs hand to the differentiable part (MPV.py:353-405)
# ------------------------------------------------------------------------------------------------
def raster_tables(st: MO.MPVState, H, W, tar_extrin, tar_intrin, device):
    """Slot-compacted masks and uv lists in the reference's flattening order (pixel-major, slot-minor)."""
    geo = MO.geometry(st, H, W, tar_extrin, tar_intrin)
    D = st.mpi_d
    order = torch.arange(D) if geo["forward_order"] else torch.arange(D - 1, -1, -1)
    hit, kind = geo["hit"][:, order], geo["kind"][:, order]
    ax, ay = geo["ax"][:, order], geo["ay"][:, order]
    P = H * W
    K = int(hit.sum(1).max().item())
    slot = torch.cumsum(hit.long(), 1) - 1
    slot_kind = torch.zeros(P, K, dtype=torch.long)
    slot_ax = torch.zeros(P, K, dtype=ax.dtype)
    slot_ay = torch.zeros(P, K, dtype=ay.dtype)
    pi = torch.arange(P)[:, None].expand(P, D)
    slot_kind[pi[hit], slot[hit]] = kind[hit]
    slot_ax[pi[hit], slot[hit]] = ax[hit]
    slot_ay[pi[hit], slot[hit]] = ay[hit]
    out = {"K": K}
    for name, k, atl in (("static", 1, st.atlas), ("dyn", 2, st.atlas_dyn)):
        m = (slot_kind == k).reshape(-1)
        hA, wA = atl.shape[-2:]
        u = slot_ax.reshape(-1)[m] / max(wA - 1, 1) * 2 - 1
        v = slot_ay.reshape(-1)[m] / max(hA - 1, 1) * 2 - 1
        out["mask_" + name] = m.to(device)
        out["uv_" + name] = torch.stack([u, v], -1).float().to(device)
    return out


# ------------------------------------------------------------------------------------------------
# differentiable part, the reference's operators
# ------------------------------------------------------------------------------------------------
def _sample_rgba(atlas_, uv):
    b, c = atlas_.shape[:2]
    feat = F.grid_sample(atlas_, uv[None, None].expand(b, 1, -1, 2), padding_mode="zeros", align_corners=True)
    feat = feat.reshape(b, c, -1).permute(0, 2, 1)
    return torch.cat([torch.sigmoid(feat[..., :-1]), torch.sigmoid(feat[..., -1:])], -1)


def render_ops(atlas, atlas_dyn, tabs, ts, H, W):
    T, K = len(ts), tabs["K"]
    rgba_s = _sample_rgba(atlas, tabs["uv_static"])
    rgba_d = _sample_rgba(atlas_dyn[ts], tabs["uv_dyn"])                     # index copy of the atlas (MPV.py:439)
    canvas = torch.zeros((1, H * W * K, 4), dtype=rgba_s.dtype, device=rgba_s.device)
    mpi = torch.masked_scatter(canvas, tabs["mask_static"][None, :, None], rgba_s)
    mpi = torch.masked_scatter(mpi.expand(T, -1, 4), tabs["mask_dyn"][None, :, None], rgba_d)
    mpi = mpi.reshape(T, H, W, K, 4)
    alpha, content = mpi[..., -1], mpi[..., :-1]
    bw = torch.cumprod((-alpha + 1)[..., :-1], dim=-1)
    bw = torch.cat([alpha[..., :1], alpha[..., 1:] * bw], dim=-1)
    rgb = (content * bw.unsqueeze(-1)).sum(dim=-2)
    return rgb, mpi


def _unfold3d(x, pt, p, st, s):
    return LL.unfold3d(x, pt, p, p, st, s, s)


def _distances(X, Y):
    X = X.reshape(*X.shape[:2], -1)
    Y = Y.reshape(*Y.shape[:2], -1)
    dist = (X * X).sum(-1)[:, :, None] + (Y * Y).sum(-1)[:, None, :] - 2.0 * (X @ Y.permute(0, 2, 1))
    dist /= X.shape[-1]
    return dist


def _nn_lowmem(X, Y, alpha, chunk=1024):
    nns = torch.zeros(X.shape[:2], dtype=torch.long, device=X.device)
    norm = 1
    if alpha is not None:
        mins = torch.zeros(Y.shape[:2], dtype=X.dtype, device=X.device)
        for a in range(0, Y.shape[1], chunk):
            mins[:, a:a + chunk] = _distances(X, Y[:, a:a + chunk]).min(1)[0]
        norm = alpha + mins[:, None]
    for a in range(0, X.shape[1], chunk):
        nns[:, a:a + chunk] = torch.argmin(_distances(X[:, a:a + chunk], Y) / norm, dim=2)
    return nns


_FOLD_INDEX = {}


def _fold_index(shape, pt, p, st, s, device):
    key = (tuple(shape), pt, p, st, s, str(device))
    if key not in _FOLD_INDEX:
        t, h, w = shape
        lin = torch.arange(t * h * w, dtype=torch.float64, device=device).reshape(1, 1, t, h, w)
        _FOLD_INDEX[key] = _unfold3d(lin, pt, p, st, s).reshape(-1).long()
    return _FOLD_INDEX[key]


def find_nn_and_merge(x, y, p, pt, s, st, alpha):
    alpha = None if alpha > 100 else alpha
    px = _unfold3d(x, pt, p, st, s)
    b, c, d, h, w = px.shape
    B = b * h * w
    px = px.permute(0, 3, 4, 2, 1).reshape(B, -1, 3, pt, p, p)
    py = _unfold3d(y, pt, p, st, s).permute(0, 3, 4, 2, 1).reshape(B, -1, 3, pt, p, p)
    nns = _nn_lowmem(px, py, alpha)
    sel = py[torch.arange(B, device=nns.device)[:, None], nns]
    cols = sel.reshape(b, h, w, d, c).permute(0, 4, 3, 1, 2).reshape(b, 3, pt * p * p, d * h * w)
    cols = torch.cat([cols, torch.ones_like(cols[:, :1])], dim=1)             # votes + weight channel
    t_, h_, w_ = x.shape[-3:]
    idx = _fold_index((t_, h_, w_), pt, p, st, s, x.device)
    out = torch.zeros(b, 4, t_ * h_ * w_, dtype=x.dtype, device=x.device)
    out.index_add_(2, idx, cols.reshape(b, 4, -1))
    out = out.reshape(b, 4, t_, h_, w_)
    return out[:, :3], out[:, 3:].clamp_min(1e-10), nns.reshape(h, w, d)


def gpnn_lowmem_ops(x, y, macro_block=64, patch_size=7, stride=2, patcht_size=7, stridet=2, rou=0, scaling=0.2,
                    alpha=1e10, return_nn=False, **_):
    p, pt, s, st = int(patch_size), int(patcht_size), int(stride), int(stridet)
    fit = LL._fit
    mb = fit(int(macro_block), p, s)
    t, h, w = x.shape[-3:]
    h, w, t = fit(h, p, s), fit(w, p, s), fit(t, pt, st)
    x = x[..., :t, :h, :w]
    y = y[..., :h, :w]
    nn_blocks = {}
    with torch.no_grad():
        ms = mb - p + s
        y2x = torch.zeros_like(x)
        weight = torch.zeros_like(x[:, :1])
        for hs in np.arange(0, h - mb + ms, ms):
            for ws in np.arange(0, w - mb + ms, ms):
                v, c, nn = find_nn_and_merge(x[..., hs:hs + mb, ws:ws + mb], y[..., hs:hs + mb, ws:ws + mb],
                                             p, pt, s, st, alpha)
                y2x[..., hs:hs + mb, ws:ws + mb] += v
                weight[..., hs:hs + mb, ws:ws + mb] += c
                if return_nn:
                    nn_blocks[(int(hs) // s, int(ws) // s)] = nn
        y2x = y2x / weight
    loss = LL.robust_lossfun(x - y2x, rou, scaling).mean()
    if return_nn:
        ho, wo = (h - p) // s + 1, (w - p) // s + 1
        full = torch.zeros(ho, wo, next(iter(nn_blocks.values())).shape[-1], dtype=torch.long, device=x.device)
        for (i, j), nn in nn_blocks.items():
            full[i:i + nn.shape[0], j:j + nn.shape[1]] = nn
        return loss, full
    return loss


def forward_train_ops(atlas, atlas_dyn, tabs, h, w, res, losscfg, mpi_d, *, isloop=True, scale_invariant=True,
                      swd_patcht_size=3, rgb_smooth_w=0.2, a_smooth_w=0.2, return_nn=False):
    """MPV.py:477-553 training branch + the weighted total of train_3dvid.py:230-240.  `res` (1,F,3,h,w)."""
    T = atlas_dyn.shape[0]
    ts = torch.arange(T, device=atlas_dyn.device)
    rgb, mpi = render_ops(atlas, atlas_dyn, tabs, ts, h, w)
    rgb = rgb.permute(0, 3, 1, 2)
    cfg = dict(losscfg)
    name = cfg.pop("loss_name")
    assert name == "gpnn_lm", "the timed baseline is the configured stage-2 loss (configs/mpv_base.txt)"
    gain = float(cfg.pop("loss_gain", 1.0))
    for k in ("dist_fn", "factor"):
        cfg.pop(k, None)
    rgb_pad = torch.cat([rgb, rgb[:swd_patcht_size - 1]], 0) if isloop else rgb
    if scale_invariant:
        res_avg = res[0].mean(dim=0)
        rgb_avg = rgb.detach().mean(dim=0)
        scale = torch.exp(torch.log((res_avg + 0.01) / (rgb_avg + 0.01)).mean())
        rgb_pad = rgb_pad * ((scale + 3) / 4)
    main = gpnn_lowmem_ops(rgb_pad.permute(1, 0, 2, 3)[None], res.permute(0, 2, 1, 3, 4), return_nn=return_nn, **cfg)
    nn = None
    if return_nn:
        main, nn = main
    extra = {"swd": main.reshape(1, -1) * gain}
    K = mpi.shape[-2]
    if rgb_smooth_w > 0:
        sm = mpi[..., :-1]
        extra["rgb_smooth"] = ((sm[:, :, :-1] - sm[:, :, 1:]).abs().mean() +
                               (sm[:, :-1] - sm[:, 1:]).abs().mean()).reshape(1, -1) * (gain * K / mpi_d)
    if a_smooth_w > 0:
        sm = mpi[..., -1]
        extra["a_smooth"] = ((sm[:, :, :-1] - sm[:, :, 1:]).abs().mean() +
                             (sm[:, :-1] - sm[:, 1:]).abs().mean()).reshape(1, -1) * (gain * K / mpi_d)
    total = extra["swd"].mean()
    if rgb_smooth_w > 0:
        total = total + extra["rgb_smooth"].mean() * rgb_smooth_w
    if a_smooth_w > 0:
        total = total + extra["a_smooth"].mean() * a_smooth_w
    return total, extra, dict(rgb=rgb, nn=nn)
