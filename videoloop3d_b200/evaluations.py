"""Evaluation-side consumers of the hot-path kernels (SURVEY.md §8(f) "next" rows N1 / N3).

* `compute_nnerr` — drop-in for `evaluations/NNMSE.py:7-58` (called 9x per view by
  `scripts/script_evaluate_ours.py:201-243`): nearest-neighbour patch error between two videos.  The NN search
  is `vl3d_patchnn_search` (alpha=None), the per-patch L1 error `vl3d_patch_l1`; only the reference's
  macro-block bookkeeping (mean over blocks of per-block means) stays on the host.
* `to8b` — `utils.py:17` on device for rendered frames (`scripts/script_render_video.py:137-149`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, ops
from .loop_loss import _planar


def compute_nnerr(src, tar, patch_size=7, stride=2, patcht_size=7, stridet=2, macro_block=65):
    """src, tar: (1,3,f,h,w) CUDA tensors.  Returns a python float, like the reference."""
    p, pt, s, st = int(patch_size), int(patcht_size), int(stride), int(stridet)
    x, y = _planar(src.detach()), _planar(tar.detach())
    if not x.is_cuda:
        raise _lib.Vl3dError("compute_nnerr runs on CUDA only")
    desc = ops.make_loss_desc(x.shape, (x.stride(0), x.stride(1), x.stride(2)), y.shape,
                              (y.stride(0), y.stride(1), y.stride(2)), p, pt, s, st, 1e10)
    nn = ops.patchnn_search(desc, x, None, y)
    err = torch.empty((desc.ho, desc.wo, desc.n1), dtype=torch.float32, device=x.device)
    _lib.call("vl3d_patch_l1", C.byref(desc), _lib.ptr(x), _lib.ptr(y), _lib.ptr(nn), _lib.ptr(err), _lib.stream_ptr())
    # macro blocks (NNMSE.py:23,33-56): block starts every macro_block - p + s pixels, mean over blocks of block means
    mb = ops._fit(int(macro_block), p, s, "macro_block")
    ms = mb - p + s
    means = []
    for hs in np.arange(0, desc.h - mb + ms, ms):
        r0, nr = int(hs) // s, (min(mb, desc.h - int(hs)) - p) // s + 1
        for ws in np.arange(0, desc.w - mb + ms, ms):
            c0, nc = int(ws) // s, (min(mb, desc.w - int(ws)) - p) // s + 1
            means.append(err[r0:r0 + nr, c0:c0 + nc].mean())
    return float(torch.stack(means).mean())


def to8b(rgb_tchw):
    """(T,3,H,W) float CUDA tensor -> (T,H,W,3) uint8, `(255*clip(x,0,1)).astype(uint8)` (utils.py:17)."""
    x = rgb_tchw.detach()
    if not x.is_contiguous():
        x = x.contiguous()
    T, c, H, W = x.shape
    assert c == 3 and x.is_cuda and x.dtype == torch.float32
    out = torch.empty((T, H, W, 3), dtype=torch.uint8, device=x.device)
    _lib.call("vl3d_to8b", _lib.ptr(x), _lib.ptr(out), int(T), int(H), int(W), _lib.stream_ptr())
    return out
