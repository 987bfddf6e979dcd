"""GPU: the fused backward + Adam kernel (`vl3d_fused_bwd_adam`) against the separate kernels
(`vl3d_composite_bwd` + `vl3d_adam_step`) and against the CPU oracle's step.

* with lr = 0 the first moment after one step is (1 - beta1) * gradient: the fused kernel's texel gradient is compared
  with the classic backward's element-wise, for every schedule (generic / band / band-zero), dense and sparse layouts,
  even and odd frame counts, oblique views (mixed tiles), image sizes that are not multiples of the tiles;
* whole optimisation steps (FusedLoopStep) with every schedule give the same parameters / optimiser state / losses as
  the separate kernels;
* the gradient buffer is all-zero again after a step with the re-zeroing schedules;
* one fused step equals the oracle's step (autograd + Adam in float64 on the CPU).
"""
import numpy as np
import pytest
import torch

from oracle import mpv_oracle as MO
from test_gpu_parity import state_tensors
from test_gpu_tma import VIEWS, _rot

pytestmark = pytest.mark.gpu

CFG = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=7, patcht_size=3, stride=4, stridet=1, alpha=0.0, rou="-2",
           scaling=0.1, dist_fn="mse", macro_block=65, factor=1)


def _view(vname, W, H, st):
    v = VIEWS[vname]
    ext = torch.eye(4)
    ext[:3, :3] = _rot(*v["rot"])
    ext[:3, 3] = torch.tensor(v["trans"])
    f = 0.8 * W * v["fmul"]
    intr = torch.tensor([[f, 0, W / 2 + 0.37], [0, f, H / 2 - 0.21], [0, 0, 1.]])
    return ext[None], intr[None]


def _dense(H, W, D, T, seed=11):
    st = MO.dense_state(H, W, D, 5, 8, 2, T, 1.0, 10.0, 1.15, 1.15, seed=seed)
    st.atlas = st.atlas[:, :, :1, :1].clone()
    return st


def _sparse(H, W, D, T, seed=4):
    return MO.sparse_state(H, W, D, 6, 9, T, 1.0, 10.0, tile=6, occupancy=0.7, dyn_frac=0.5, h_scale=1.2, w_scale=1.2, seed=seed)


def _target(F, H, W, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, F, 3, H, W, generator=g).to(dev)


def _run_steps(st, H, W, mode, vname, n_steps, lr, dev, opts=None, smooth=True):
    from videoloop3d_b200 import FusedLoopStep
    from videoloop3d_b200.testing import model_from_tensors
    kw = {} if smooth else dict(rgb_smooth_loss_weight=0.0, a_smooth_loss_weight=0.0)
    m = model_from_tensors(state_tensors(st), H, W, dev, **kw)
    step = FusedLoopStep(m, fused=mode, fused_opts=opts)
    ext, intr = _view(vname, W, H, st)
    res = _target(9, H, W, dev)
    outs = []
    for _ in range(n_steps):
        outs.append(step.step(H, W, ext, intr, res, CFG, lr=lr))
    torch.cuda.synchronize()
    return m, step, outs


@pytest.mark.parametrize("layout", ["dense", "sparse"])
@pytest.mark.parametrize("vname", ["near_identity", "oblique", "roll", "zoom_out"])
@pytest.mark.parametrize("T", [4, 7])
def test_fused_gradient_equals_classic_backward(layout, vname, T):
    dev = torch.device("cuda:0")
    H, W, D = 75, 133, 6
    st = _dense(H, W, D, T) if layout == "dense" else _sparse(H, W, D, T)
    m0, s0, o0 = _run_steps(st, H, W, "off", vname, 1, 0.0, dev)
    g_ref = s0._state["atlas_dyn"][0] / 0.1                         # m = (1 - beta1) * g after the first step
    scale = float(g_ref.abs().max())
    assert scale > 0
    modes = ["generic"] + (["band", "band-zero", "own"] if layout == "dense" else [])
    for mode in modes:
        for opts in ({}, dict(row_block=3, zero_ahead=1, adam_lag=0, ctas_per_sm=1)):
            m1, s1, o1 = _run_steps(st, H, W, mode, vname, 1, 0.0, dev, opts)
            assert s1.last_schedule is not None and s1.last_schedule.kind == mode
            g = s1._state["atlas_dyn"][0] / 0.1
            err = float((g - g_ref).abs().max())
            assert err <= 3e-6 * scale, (mode, opts, err, scale)     # same terms; RED order differs
            assert torch.equal(m1.atlas_dyn.data, m0.atlas_dyn.data)  # lr = 0
            for k in o0[0]:
                assert abs(float(o1[0][k]) - float(o0[0][k])) <= 1e-6 * abs(float(o0[0][k])) + 1e-9, (mode, k)
            if mode != "band-zero":
                assert float(s1._buf["g_dyn"].abs().max()) == 0.0     # re-zeroed by the Adam items


@pytest.mark.parametrize("layout,mode", [("dense", "generic"), ("dense", "band"), ("dense", "band-zero"), ("dense", "own"),
                                         ("sparse", "generic")])
def test_fused_steps_equal_separate_kernels(layout, mode):
    dev = torch.device("cuda:0")
    H, W, D, T = 64, 96, 8, 6
    st = _dense(H, W, D, T, seed=3) if layout == "dense" else _sparse(H, W, D, T, seed=5)
    m0, s0, o0 = _run_steps(st, H, W, "off", "near_identity", 3, 0.01, dev)
    m1, s1, o1 = _run_steps(st, H, W, mode, "near_identity", 3, 0.01, dev)
    for a, b in zip(o0, o1):
        for k in a:
            assert abs(float(a[k]) - float(b[k])) <= 2e-5 * abs(float(a[k])) + 1e-8, k
    # Adam's eps = 6e-8 turns a sign flip of a ~0 gradient into a 2*lr difference: bound the count, not the maximum
    d = (m1.atlas_dyn.data - m0.atlas_dyn.data).abs()
    assert float((d > 1e-5).float().mean()) < 2e-4, float((d > 1e-5).float().mean())
    assert float(d.median()) < 1e-7
    dm = (s1._state["atlas_dyn"][0] - s0._state["atlas_dyn"][0]).abs().max()
    assert float(dm) <= 1e-5 * float(s0._state["atlas_dyn"][0].abs().max())
    if layout == "sparse":
        assert float((m1.atlas.data - m0.atlas.data).abs().max()) < 0.021


def test_fused_without_regulariser():
    dev = torch.device("cuda:0")
    H, W, D, T = 64, 96, 8, 4
    st = _sparse(H, W, D, T, seed=7)
    _, s0, o0 = _run_steps(st, H, W, "off", "near_identity", 1, 0.0, dev, smooth=False)
    _, s1, o1 = _run_steps(st, H, W, "generic", "near_identity", 1, 0.0, dev, smooth=False)
    g0, g1 = s0._state["atlas_dyn"][0], s1._state["atlas_dyn"][0]
    assert float((g1 - g0).abs().max()) <= 3e-6 * float(g0.abs().max())
    assert abs(float(o1[0]["loss"]) - float(o0[0]["loss"])) <= 1e-6 * abs(float(o0[0]["loss"]))


def test_fused_step_matches_oracle_step():
    """One fused (band-zero) step on the dense layout vs the oracle: loss 1e-4 relative, updated texels: Adam's first
    step moves every texel by lr * sign(g) (|g| >> eps), so the update direction must agree wherever the oracle's
    gradient is not tiny."""
    dev = torch.device("cuda:0")
    H, W, D, T = 48, 80, 6, 4
    st = _dense(H, W, D, T, seed=21)
    lr = 0.01
    m1, s1, o1 = _run_steps(st, H, W, "band-zero", "near_identity", 1, lr, dev)
    ext, intr = _view("near_identity", W, H, st)
    res = _target(9, H, W, torch.device("cpu"))
    ad = st.atlas_dyn.double().requires_grad_(True)
    a = st.atlas.double().requires_grad_(True)
    extra, _ = MO.forward_train(st, H, W, ext, intr, res, CFG, atlas=a, atlas_dyn=ad)
    loss = MO.total_loss(extra)
    loss.backward()
    assert abs(float(o1[0]["loss"]) - float(loss)) <= 1e-4 * abs(float(loss))
    p_ref, _, _ = MO.adam_step(ad.detach(), ad.grad, torch.zeros_like(ad), torch.zeros_like(ad), 1, lr)
    upd = (m1.atlas_dyn.data.cpu().double() - st.atlas_dyn.double())
    upd_ref = p_ref - st.atlas_dyn.double()
    big = ad.grad.abs() > 1e-4 * ad.grad.abs().max()
    # (the regulariser's sign() terms may flip where two neighbouring activations agree to ~1e-6: a bounded handful)
    bad = (upd[big] - upd_ref[big]).abs() > 1e-3 * lr
    assert float(bad.float().mean()) < 1e-3, float(bad.float().mean())
    assert float((upd - upd_ref).abs().mean()) <= 1e-2 * lr


def test_compressible_gradient_buffer_is_a_placement_hint(monkeypatch):
    """ops.compressible_zeros_like: same shape / strides / zeros as zeros_like (or None where the device does not offer
    compressible memory), and a step gives bit-identical parameters with and without it."""
    from videoloop3d_b200 import ops
    dev = torch.device("cuda:0")
    ref = torch.empty((3, 4, 17, 23), device=dev).contiguous(memory_format=torch.channels_last)
    t, granted = ops.compressible_zeros_like(ref)
    if t is not None:
        assert t.shape == ref.shape and t.stride() == ref.stride() and t.device == ref.device
        assert float(t.abs().max()) == 0.0
        t[1, 2, 3, 4] = 5.0
        assert float(t.sum()) == 5.0
        del t
    H, W, D, T = 64, 96, 8, 4
    st = _dense(H, W, D, T, seed=3)
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("VL3D_GRAD_COMPRESS", flag)
        m, s_, o = _run_steps(st, H, W, "generic", "near_identity", 2, 0.01, dev)
        outs[flag] = (m.atlas_dyn.data.clone(), float(o[-1]["loss"]), s_.grad_compressed)
    assert outs["0"][2] is False
    assert abs(outs["1"][1] - outs["0"][1]) <= 1e-6 * abs(outs["0"][1])
    d = (outs["1"][0] - outs["0"][0]).abs()
    assert float((d > 1e-5).float().mean()) < 2e-4 and float(d.median()) < 1e-7    # (RED order is not deterministic)
