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


def _tie_budget(*slot_tensors, thr=1e-5):
    """Neighbouring slot values that tie to within the fp32 sampling noise: the gradient of |a - b| is undefined there; each
    such pair may move the 2 x 4 texels it taps (see tests/test_gpu_stage1.py)."""
    n = 0
    for t in slot_tensors:
        for d in ((t[:, :, :-1] - t[:, :, 1:]).abs(), (t[:, :-1] - t[:, 1:]).abs()):
            n += int(((d > 0) & (d < thr)).sum())
    return 8 * n


def test_stage1_step_shape_matches_oracle():
    """The stage-1 model at configs/mpi_base.txt's shape (D = 32, 36 x 64 vertices, scale 1.6, 180 x 320 patch; loop mask,
    every extra term on) — forward values and the gradients of atlas / atlas_mask against the oracle's restatement of
    MPI.py:452-652 (which tests/test_oracle.py pins to the unmodified reference)."""
    from videoloop3d_b200 import MPMesh, default_args_stage1
    dev = torch.device("cuda:0")
    hv, wv = 36, 64
    st, atlas_mask = MO.stage1_state(H, W, D, hv, wv, 4, 1.0, 10.0, 1.6, 1.6, seed=31, alpha_mean=-2.5)
    ext, intr = _view()
    weights = dict(sparsity=0.004, rgb_smooth=0.2, a_smooth=0.5, density=0.02, d_smooth=0.1, l_smooth=0.05)
    a = st.atlas.double().requires_grad_(True)
    am = atlas_mask.double().requires_grad_(True)
    rgbl_o, extra_o, var_o = MO.forward_stage1(st, H, W, ext, intr, 1.0, 10.0, edge_scale=0.5, atlas=a, atlas_mask=am)
    assert var_o["K"] == D
    g_up = torch.rand(rgbl_o.shape, generator=torch.Generator().manual_seed(7), dtype=torch.float64) - 0.4
    ((rgbl_o * g_up).mean() + sum(extra_o[k] * w for k, w in weights.items())).backward()
    args = default_args_stage1(edge_scale=0.5, **{k + "_loss_weight": w for k, w in weights.items()})
    f = 0.8 * W
    m = MPMesh(args, H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32),
               1.0, 10.0)
    assert torch.equal(m.faces, st.faces) and torch.allclose(m.uvs.data, st.uvs, atol=1e-7) and tuple(m.atlas.shape) == tuple(st.atlas.shape)
    m.atlas.data, m.atlas_mask.data = st.atlas.clone(), atlas_mask.clone()
    m = m.to(dev).train()
    rgbl, extra = m(H, W, ext.to(dev), intr.to(dev))
    assert float((rgbl.detach().cpu().double() - rgbl_o.detach()).abs().max()) < 1e-4
    for k in weights:
        assert abs(float(extra[k]) - float(extra_o[k])) < 1e-4 * max(abs(float(extra_o[k])), 1e-3), k
    ((rgbl * g_up.to(dev).float()).mean() + sum(extra[k].mean() * w for k, w in weights.items())).backward()
    budget = _tie_budget(var_o["mpi"].detach(), var_o["loopmask3d"].detach())
    for name, got, ref in (("atlas", m.atlas.grad, a.grad), ("atlas_mask", m.atlas_mask.grad, am.grad)):
        err = (got.cpu().double() - ref).abs()
        bad = int((err > 5e-4 * float(ref.abs().max())).sum())
        assert bad <= budget, (name, bad, budget, float(err.max()))
        assert float(err.max()) < 0.5 * float(ref.abs().max()), (name, float(err.max()))


def test_optional_terms_at_step_shape_match_oracle():
    """alpha / disparity / sparsity of the DENSE stage-2 model at the reference's step shape (D = 32, 180 x 320, 4 frames)
    and the gradients of a random functional of them (csrc/terms.cu) against the oracle."""
    from videoloop3d_b200.testing import model_from_tensors
    dev = torch.device("cuda:0")
    Tt = 4
    st = MO.dense_state(H, W, D, 27, 48, 4, Tt, 1.0, 10.0, 1.1, 1.1, seed=33, alpha_mean=-2.5)
    st.atlas = st.atlas[:, :, :1, :1].clone()
    ext, intr = _view()
    ts = list(range(Tt))
    ad = st.atlas_dyn.double().requires_grad_(True)
    _, var_o = MO.render(st, H, W, ext, intr, ts, atlas_dyn=ad)
    al = var_o["mpi"][..., -1]
    sp_o = (al.norm(dim=-1, p=1) / al.norm(dim=-1, p=2).clamp_min(1e-4)).sum()
    gen = torch.Generator().manual_seed(9)
    ga = torch.rand(var_o["alpha"].shape, generator=gen, dtype=torch.float64) - 0.3
    gd = torch.rand(var_o["alpha"].shape, generator=gen, dtype=torch.float64) - 0.5
    ((var_o["alpha"] * ga).sum() + (var_o["disp_norm"] * gd).sum() + 0.37 * sp_o).backward()
    m = model_from_tensors(state_tensors(st), H, W, dev)
    extrin = ext @ torch.inverse(st.ref_extrin)[None]
    view = m.make_view(H, W, extrin, intr)
    alpha, disp, sp = m._render_terms(view, m._ts_tensor(ts), Tt, H, W, extrin, intr, want_disp=True, want_sparsity=True)
    assert float((alpha.detach().cpu().double() - var_o["alpha"].detach()).abs().max()) < 1e-4
    assert float((disp.detach().cpu().double() - var_o["disp_norm"].detach()).abs().max()) < 1e-4 * float(var_o["disp_norm"].abs().max())
    assert abs(float(sp) - float(sp_o)) < 1e-4 * float(sp_o)
    ((alpha * ga.to(dev).float()).sum() + (disp * gd.to(dev).float()).sum() + 0.37 * sp.sum()).backward()
    err = float((m.atlas_dyn.grad.cpu().double() - ad.grad).abs().max())
    assert err < 3e-4 * float(ad.grad.abs().max()), err
    assert float(m.atlas_dyn.grad[:, :3].abs().max()) == 0.0
