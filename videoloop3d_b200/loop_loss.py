"""Drop-in looping-loss callables (reference: utils_vid.py:265-445, registered at MPV.py:131-138).

Same call signature, kwargs and side-effect caches (`last_y2x`, `last_weight`) as the reference
objects; the work is done by `vl3d_patchnn_search` + `vl3d_vote_loss` (CUDA).  The macro-block loop of
the reference's low-memory variant only bounds the size of its im2col / distance temporaries and does
not change the result (README.md:145), so `macro_block` is accepted and ignored: the fused kernels
never build those temporaries.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import Vl3dError


def _planar(v):
    """(1,3,t,h,w) tensor -> (t,3,h,w) view with unit pixel stride (copy only if needed)."""
    if v.dim() != 5 or v.shape[0] != 1 or v.shape[1] != 3:
        raise Vl3dError(f"expected a (1,3,t,h,w) video, got {tuple(v.shape)} (batches not implemented, as in the reference)")
    p = v[0].permute(1, 0, 2, 3)
    if p.stride(3) != 1 and p.shape[3] != 1:
        p = p.contiguous()
    return p


class _GPNNLoss:
    fit = True

    def __init__(self):
        self.last_y2x = None
        self.last_weight = None
        self.last_nn = None

    def _run(self, x, y, same_input, cfg, xscale=None):
        xp, yp = _planar(x), _planar(y.detach())
        if not xp.is_cuda:
            raise Vl3dError("the looping loss runs on CUDA only (no CPU fallback)")
        return _LoopLossCached.apply(xp, xscale, yp, cfg, self, bool(same_input), self.fit)

    def planar(self, x_tchw, xscale, y_fchw, cfg):
        """Fast path used by MPMeshVid.forward: x (Tx,3,h,w), y (F,3,h,w) already planar."""
        return _LoopLossCached.apply(x_tchw, xscale, y_fchw, cfg, self, False, self.fit)


class _LoopLossCached(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, xscale, y, cfg, owner, same_input, fit):
        sx = (x.stride(0), x.stride(1), x.stride(2))
        sy = (y.stride(0), y.stride(1), y.stride(2))
        desc = ops.make_loss_desc(x.shape, sx, y.shape, sy, cfg["patch_size"], cfg["patcht_size"], cfg["stride"],
                                  cfg["stridet"], cfg.get("alpha", 1e10), fit=fit)
        if same_input and owner.last_nn is not None:            # utils_vid.py:300-302
            nn = owner.last_nn
        else:
            nn = ops.patchnn_search(desc, x, xscale, y)
        loss, grad, y2x, wgt = ops.vote_loss(desc, x, xscale, y, nn, cfg.get("rou", 0), cfg.get("scaling", 0.2), 1.0,
                                             (x.shape[0], x.shape[2], x.shape[3]), want_cache=True)
        owner.last_y2x, owner.last_weight, owner.last_nn = y2x, wgt, nn
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None


class Patch3DGPNNLowMemLoss(_GPNNLoss):
    """utils_vid.py:289-349."""
    fit = True

    def __call__(self, x, y, mask=None, same_input=False, macro_block=64, patch_size=7, stride=2, patcht_size=7,
                 stridet=2, rou=0, scaling=0.2, **kwargs):
        cfg = dict(patch_size=patch_size, stride=stride, patcht_size=patcht_size, stridet=stridet, rou=rou,
                   scaling=scaling, alpha=kwargs.get("alpha", 1e10))
        _check_dist(kwargs)
        return self._run(x, y, same_input, cfg)


class Patch3DGPNNDirectLoss(_GPNNLoss):
    """utils_vid.py:265-286 (no size fitting: uncovered border pixels keep y2x = 0, weight = 1e-10)."""
    fit = False

    def __call__(self, x, y, mask=None, same_input=False, rou=0, scaling=0.2, patch_size=7, patcht_size=7,
                 stride=1, stridet=1, **kwargs):
        cfg = dict(patch_size=patch_size, stride=stride, patcht_size=patcht_size, stridet=stridet, rou=rou,
                   scaling=scaling, alpha=kwargs.get("alpha", 1e10))
        _check_dist(kwargs)
        return self._run(x, y, same_input, cfg)


def _check_dist(kwargs):
    if kwargs.get("dist_fn", "mse") != "mse":
        raise NotImplementedError("dist_fn='ssim' is never configured by the reference and is out of scope")


class _VideoLoss(torch.autograd.Function):
    """Patch3DMSE / Patch3DAvg on (1,3,t,h,w) videos through `vl3d_video_loss` (value + dL/dx in one launch)."""

    @staticmethod
    def forward(ctx, x, y, kind):
        for v in (x, y):
            if v.dim() != 5 or v.shape[0] != 1 or v.shape[1] != 3:
                raise Vl3dError(f"expected a (1,3,t,h,w) video, got {tuple(v.shape)}")
        if not x.is_cuda:
            raise Vl3dError("the looping losses run on CUDA only (no CPU fallback)")
        if tuple(x.shape[-2:]) != tuple(y.shape[-2:]):
            raise ValueError("x and y must have the same spatial size")
        xc, yc = x[0].contiguous().float(), y[0].detach().contiguous().float()
        grad = torch.empty_like(xc) if x.requires_grad else None
        lib = ops._lib
        part = torch.empty(lib.load().vl3d_video_loss_partials(), dtype=torch.float64, device=x.device)
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        lib.call("vl3d_video_loss", int(kind), lib.ptr(xc), lib.ptr(yc), int(xc.shape[1]), int(yc.shape[1]), int(xc.shape[2]),
                 int(xc.shape[3]), lib.ptr(grad), lib.ptr(part), lib.ptr(out), lib.stream_ptr())
        ctx.save_for_backward(grad)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (None if grad is None else (grad * g)[None]), None, None


def Patch3DMSE(x, y, **kwargs):
    """utils_vid.py:437-440: mean squared difference over the frames both videos have."""
    return _VideoLoss.apply(x, y, 0)


def Patch3DAvg(x, y, **kwargs):
    """utils_vid.py:443-445: mean squared difference of the temporal means."""
    return _VideoLoss.apply(x, y, 1)
