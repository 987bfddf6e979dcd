"""GPU: the optimisation step at the reference's OWN step shape against the CPU oracle — 180x320 patch, D = 32 planes,
T = 48 frames, F = 66 target frames (configs/mpv_base.txt:14,21-24), both loss configurations of train_3dvid.py:161-190
and both atlas layouts:

* dense layout + reference-view loss (p=11, pt=3, s=4, alpha=0): the product instantiations — TMA render,
  the fused backward + Adam kernel with TMA-staged tiles (`fused_bwd_adam_kernel<true,3>` = `bwd_tile<2,true,3>`),
  `patchnn_strip8_kernel<2,3,true>` — at real tile counts;
* tile-culled (sparse) layout with static + dynamic tiles + other-view loss (p=3, pt=3, s=2, alpha=None):
  per-thread loads, the generic fused kernel, the 4x4 strip search, static-atlas Adam.

Checked: rendered frames (1e-4), every loss term (1e-4 relative), NN index map (bit-exact up to proven fp64 near-ties),
texel gradients (through Adam's first moment with lr = 0; regulariser sign flips bounded as in test_gpu_parity).
The oracle renders in float32 / searches in float64 and needs ~10-40 s of host time per case."""
import numpy as np
import pytest
import torch

from oracle import looploss_oracle as LL
from oracle import mpv_oracle as MO
from test_gpu_parity import state_tensors

pytestmark = pytest.mark.gpu

H, W, D, T, F = 180, 320, 32, 48, 66
REF_CFG = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=11, patcht_size=3, stride=4, stridet=1, alpha=0.0, rou="-2",
               scaling=0.1, dist_fn="mse", macro_block=65, factor=1)
OTHER_CFG = dict(loss_name="gpnn_lm", loss_gain=1.0, patch_size=3, patcht_size=3, stride=2, stridet=1, alpha=10000.0,
                 rou="-2", scaling=0.1, dist_fn="mse", macro_block=65, factor=1)


def _view():
    ang = 0.05
    ext = torch.eye(4)
    ext[:3, :3] = torch.tensor([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]).float()
    ext[:3, 3] = torch.tensor([0.08, -0.03, 0.02])
    f = 0.8 * W
    intr = torch.tensor([[f, 0, W / 2 + 0.21], [0, f, H / 2 - 0.13], [0, 0, 1.]])
    return ext[None], intr[None]


def _target(seed):
    g = torch.Generator().manual_seed(seed)
    res = torch.rand(1, F, 3, H, W, generator=g)
    return (res + res.roll(1, 1) + res.roll(1, 3) + res.roll(1, 4)) / 4


@pytest.mark.parametrize("layout,cfg", [("dense", REF_CFG), ("sparse", OTHER_CFG)], ids=["dense-refcfg", "sparse-othercfg"])
def test_reference_step_shape_matches_oracle(layout, cfg):
    from videoloop3d_b200 import FusedLoopStep
    from videoloop3d_b200.testing import model_from_tensors
    dev = torch.device("cuda:0")
    if layout == "dense":
        st = MO.dense_state(H, W, D, 9, 16, 4, T, 1.0, 10.0, 1.0, 1.0, seed=2)
        st.atlas = st.atlas[:, :, :1, :1].clone()
    else:
        st = MO.sparse_state(H, W, D, 12, 21, T, 1.0, 10.0, tile=16, occupancy=0.25, dyn_frac=0.5, h_scale=1.1, w_scale=1.1,
                             seed=3)
    ext, intr = _view()
    res = _target(7)
    m = model_from_tensors(state_tensors(st), H, W, dev)
    step = FusedLoopStep(m)
    out = step.step(H, W, ext, intr, res.to(dev), cfg, lr=0.0)
    torch.cuda.synchronize()
    pack = m.mesh_pack()
    assert pack.rect_planes == (layout == "dense")
    assert step.last_schedule is not None and step.last_schedule.kind == "generic"   # ("auto" at this shape / chunk count)

    # NN search on IDENTICAL inputs: the oracle's float64 search over the video the CUDA path searched (its own render
    # times its own gain) must select the same indices, except where two candidates tie to 1e-5 in float64
    x_cuda = step._buf["x_scaled"].cpu().double().permute(1, 0, 2, 3)[None]
    y64 = res.permute(0, 2, 1, 3, 4).double()
    lcfg = {k: cfg[k] for k in ("patch_size", "patcht_size", "stride", "stridet", "alpha", "rou", "scaling")}
    _, aux_nn = LL.gpnn_lowmem(x_cuda, y64, macro_block=65, **lcfg)
    nn_c = step._buf["nn"].cpu().long()
    assert nn_c.shape == aux_nn["nn"].shape
    mism = int((nn_c != aux_nn["nn"]).sum())
    if mism:
        ycrop = y64[..., :aux_nn["x"].shape[-2], :aux_nn["x"].shape[-1]]
        mism, bad = LL.tie_margin_ok(aux_nn["x"].detach(), ycrop, aux_nn["nn"], nn_c, cfg["patch_size"], cfg["patcht_size"],
                                     cfg["stride"], cfg["stridet"], cfg["alpha"])
        assert bad == 0, f"{bad} NN mismatches that are not fp64 near-ties"
        assert mism <= 1e-5 * nn_c.numel() + 2, f"{mism} near-tie NN mismatches out of {nn_c.numel()}"
    # the whole step on the oracle (float32 render), continuing with the CUDA path's (equally valid) matches
    a = st.atlas.float().requires_grad_(True)
    ad = st.atlas_dyn.float().requires_grad_(True)
    extra, aux = MO.forward_train(st, H, W, ext, intr, res, dict(cfg, nn_override=nn_c), dtype=torch.float32, atlas=a,
                                  atlas_dyn=ad)
    rgb = step._buf["rgb_pad"][:T].cpu()
    # a ray that meets the border of a plane to within rounding may see one plane more or less (the oracle decides in
    # float64, the kernel in float32): such a pixel differs in every frame; allow a handful among the 57 600
    err_px = (rgb - aux["rgb"].detach()).abs().amax(dim=(0, 1))
    bad_px = err_px > 1e-4 * float(aux["rgb"].abs().max())
    assert int(bad_px.sum()) <= 4, int(bad_px.sum())
    for k in ("swd", "rgb_smooth", "a_smooth"):
        assert abs(float(out[k]) - float(extra[k])) < 1e-4 * abs(float(extra[k])), (k, float(out[k]), float(extra[k]))
    MO.total_loss(extra).backward()
    w_max = 2 * 0.2 * cfg["loss_gain"] / (T * (H - 1) * (W - 1) * D)          # two sign flips of one pair term
    pairs = [(step._state["atlas_dyn"][0].cpu() / 0.1, ad.grad)]             # m = (1 - beta1) g after one step
    if pack.n_static:
        pairs.append((step._state["atlas"][0].cpu() / 0.1, a.grad))
    for got, ref in pairs:
        scale = float(ref.abs().max())
        assert scale > 0
        err = (got - ref).abs()
        # (regulariser sign flips move a texel by multiples of w_max; a border pixel as above moves its taps more)
        n_off = int((err > 8 * w_max + 5e-4 * scale).sum())
        assert n_off <= max(64, 2e-4 * ref.numel()), f"{n_off} texel gradients off"
    assert torch.equal(m.atlas_dyn.data.cpu(), st.atlas_dyn.float())         # lr = 0
