"""GPU parity: the CUDA path (through the C ABI) against the golden vectors of the unmodified
reference and against the CPU oracle on seeded inputs.

Tolerances (north_star): rendered RGB and loss within 1e-4 relative fp32; NN indices bit-exact (the
only admissible disagreement is an exact near-tie: fp64 margin < 1e-5 relative, SURVEY H2 — none occurs
on these inputs, which is asserted)."""
import numpy as np
import pytest
import torch

from oracle import looploss_oracle as LL
from oracle import mpv_oracle as MO
from util import cfg_from_golden, load_golden, relerr, state_from_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def dev():
    return torch.device("cuda:0")


def model_from_golden(g, **kw):
    from videoloop3d_b200.testing import model_from_tensors
    return model_from_tensors(g, int(g["H"]), int(g["W"]), dev(), **kw)


def state_tensors(st):
    return dict(verts=st.verts, planedepth=st.planedepth, faces=st.faces, faces_dyn=st.faces_dyn, uvs=st.uvs,
                uvs_dyn=st.uvs_dyn, uvfaces=st.uvfaces, uvfaces_dyn=st.uvfaces_dyn, atlas=st.atlas,
                atlas_dyn=st.atlas_dyn, ref_extrin=st.ref_extrin, ref_intrin=st.ref_intrin, mpi_d=st.mpi_d,
                hv=st.hv, wv=st.wv)


def test_library_is_loaded_and_native():
    from videoloop3d_b200 import _lib
    lib = _lib.load()
    assert lib.vl3d_version() == 100
    maps = open("/proc/self/maps").read()
    assert "libvl3d.so" in maps


@pytest.mark.parametrize("name", ["render_dense", "render_sparse"])
def test_render_matches_reference_golden(name):
    g = load_golden(name)
    m = model_from_golden(g)
    H, W, T = int(g["H"]), int(g["W"]), int(g["T"])
    ext = torch.as_tensor(g["tar_extrin"]).to(dev())
    intr = torch.as_tensor(g["tar_intrin"]).to(dev())
    m.eval()
    with torch.no_grad():
        rgb, var = m.render(H, W, ext @ m.ref_extrin[None].inverse(), intr, list(range(T)))
        assert tuple(rgb.shape) == (T, H, W, 3)
        assert relerr(rgb.cpu(), g["rgb"]) < RTOL
        assert relerr(var["alpha"].cpu(), g["alpha"]) < RTOL
        assert var["mpi"].shape[-2] == int(g["K"])                     # adaptive K (utils.py:64-69)
        assert relerr(var["mpi"].cpu(), g["mpi"]) < RTOL
        assert relerr(var["blend_weight"].cpu(), g["blend_weight"]) < RTOL
        assert torch.equal(var["pix_to_face"].cpu() >= 0, torch.as_tensor(g["pix_to_face"]) >= 0)
        rgb2, extra = m(H, W, ext, intr, ts=[T - 1, 0])
        assert extra == {} and tuple(rgb2.shape) == (2, 3, H, W)
        assert relerr(rgb2.cpu(), g["rgb_eval_ts"]) < RTOL


@pytest.mark.parametrize("name", ["step_dense_refcfg", "step_sparse_othercfg"])
def test_train_forward_backward_matches_reference_golden(name):
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    m = model_from_golden(g, rgb_smooth_loss_weight=float(g["rgb_smooth_w"]), a_smooth_loss_weight=float(g["a_smooth_w"]),
                          swd_patcht_size=int(cfg["patcht_size"]))
    H, W = int(g["H"]), int(g["W"])
    ext = torch.as_tensor(g["tar_extrin"]).to(dev())
    intr = torch.as_tensor(g["tar_intrin"]).to(dev())
    res = torch.as_tensor(g["res"]).to(dev())
    batched = {k: ([v] if isinstance(v, str) else torch.tensor([v])) for k, v in cfg.items()}
    m.train()
    none, extra = m(H, W, ext, intr, res=res, losscfg=batched)
    assert none is None
    for k in ("swd", "rgb_smooth", "a_smooth"):
        assert tuple(extra[k].shape) == (1, 1)
        assert abs(float(extra[k]) - float(g["extra_" + k])) < RTOL * abs(float(g["extra_" + k])), k
    lossobj = m.losses[cfg["loss_name"]]
    assert relerr(lossobj.last_y2x.cpu(), g["y2x"]) < 1e-5
    assert torch.equal(lossobj.last_weight.cpu(), torch.as_tensor(g["weight"]))
    loss = extra["swd"].mean() + extra["rgb_smooth"].mean() * m.args.rgb_smooth_loss_weight \
        + extra["a_smooth"].mean() * m.args.a_smooth_loss_weight
    assert abs(float(loss) - float(g["loss"])) < RTOL * abs(float(g["loss"]))
    loss.backward()
    assert m.uvs.grad is None and m.uvs_dyn.grad is None and m._verts.grad is None
    assert relerr(m.atlas_dyn.grad.cpu(), g["grad_atlas_dyn"]) < 5e-4
    if g["grad_atlas"].size > 4:
        assert relerr(m.atlas.grad.cpu(), g["grad_atlas"]) < 5e-4


@pytest.mark.parametrize("name", ["step_dense_refcfg", "step_sparse_othercfg"])
def test_adam_matches_reference_update(name):
    """Feed the reference's own gradients: the parameter update must agree to fp32 rounding."""
    from videoloop3d_b200 import FusedAdam
    g = load_golden(name)
    p = torch.nn.Parameter(torch.as_tensor(g["atlas_dyn"]).to(dev()).contiguous(memory_format=torch.channels_last))
    p.grad = torch.as_tensor(g["grad_atlas_dyn"]).to(dev())           # differently strided on purpose
    opt = FusedAdam([p], lr=float(g["lr"]), betas=(0.9, 0.999), eps=6e-8)
    opt.step()
    assert float((p.detach().cpu() - torch.as_tensor(g["new_atlas_dyn"])).abs().max()) < 1e-6
    # multi-step against torch.optim.Adam on random data, including zero gradients
    torch.manual_seed(0)
    q0 = torch.randn(3, 4, 5, 7, device=dev())
    qa, qb = torch.nn.Parameter(q0.clone()), torch.nn.Parameter(q0.clone())
    oa = FusedAdam([qa], lr=0.01, betas=(0.9, 0.999), eps=6e-8)
    ob = torch.optim.Adam([qb], lr=0.01, betas=(0.9, 0.999), eps=6e-8)
    for i in range(5):
        gr = torch.randn_like(q0) * (0.0 if i == 3 else 1e-3)
        qa.grad, qb.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    assert float((qa - qb).abs().max()) < 1e-6


@pytest.mark.parametrize("name", ["loss_lm_alpha0", "loss_lm_noalpha", "loss_direct_p7", "loss_lm_abs"])
def test_loss_matches_reference_golden(name):
    import videoloop3d_b200 as V
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    x = torch.as_tensor(g["x"]).to(dev()).requires_grad_(True)
    y = torch.as_tensor(g["y"]).to(dev())
    lossobj = getattr(V, str(g["cls"]))()
    loss = lossobj(x, y, **cfg)
    loss.backward()
    assert torch.equal(lossobj.last_nn.cpu().long(), torch.as_tensor(g["nn"]))       # bit-exact NN selection
    assert abs(float(loss) - float(g["loss"])) < RTOL * abs(float(g["loss"]))
    assert relerr(lossobj.last_y2x.cpu(), g["y2x"]) < 1e-5
    assert torch.equal(lossobj.last_weight.cpu(), torch.as_tensor(g["weight"]))
    gx = np.zeros_like(g["x"]); gref = g["grad_x"]
    assert relerr(x.grad.cpu(), gref) < RTOL
    # same_input=True re-uses the cached matches (utils_vid.py:300-302)
    loss2 = lossobj(x.detach(), y, same_input=True, **cfg)
    assert abs(float(loss2) - float(loss)) < 1e-6


CASES = [
    # D, hv, wv, H, W, T, F, kind, cfg
    dict(D=8, hv=6, wv=9, H=40, W=72, T=7, F=12, kind="dense", p=5, pt=3, s=2, st=1, alpha=0.0, rou="-2", seed=11),
    dict(D=32, hv=5, wv=7, H=37, W=45, T=5, F=9, kind="sparse", p=3, pt=3, s=2, st=1, alpha=1e4, rou="-2", seed=12),
    dict(D=5, hv=4, wv=4, H=33, W=65, T=9, F=20, kind="sparse", p=7, pt=2, s=3, st=2, alpha=0.3, rou="0", seed=13),
    dict(D=16, hv=9, wv=12, H=64, W=96, T=4, F=70, kind="dense", p=11, pt=3, s=4, st=1, alpha=0.0, rou="2", seed=14),
]


def _make_case(c):
    if c["kind"] == "dense":
        st = MO.dense_state(c["H"], c["W"], c["D"], c["hv"], c["wv"], 1, c["T"], 1.0, 10.0, 1.25, 1.25, seed=c["seed"])
        st.atlas = st.atlas[:, :, :1, :1].clone()
    else:
        st = MO.sparse_state(c["H"], c["W"], c["D"], c["hv"], c["wv"], c["T"], 1.0, 10.0, tile=5, occupancy=0.7,
                             dyn_frac=0.6, h_scale=1.25, w_scale=1.25, seed=c["seed"])
    g = torch.Generator().manual_seed(c["seed"])
    ang = 0.04
    ext = torch.eye(4)
    ext[:3, :3] = torch.tensor([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]).float()
    ext[:3, 3] = torch.tensor([0.06, -0.02, 0.03])
    f = 0.8 * c["W"]
    intr = torch.tensor([[f, 0, c["W"] / 2], [0, f, c["H"] / 2], [0, 0, 1.]])
    intr[:2, 2] += torch.rand(2, generator=g) - 0.5
    res = torch.rand(1, c["F"], 3, c["H"], c["W"], generator=g)
    res = (res + res.roll(1, 1) + res.roll(1, 3)) / 3
    cfg = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=c["p"], patcht_size=c["pt"], stride=c["s"],
               stridet=c["st"], alpha=c["alpha"], rou=c["rou"], scaling=0.1, dist_fn="mse", macro_block=65, factor=1)
    return st, ext[None], intr[None], res, cfg


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c['kind']}-D{c['D']}-p{c['p']}")
def test_fused_step_matches_oracle(case):
    """FusedLoopStep (no autograd) == autograd path == CPU oracle, including the Adam update."""
    from videoloop3d_b200 import FusedLoopStep
    from videoloop3d_b200.testing import model_from_tensors
    st, ext, intr, res, cfg = _make_case(case)
    H, W = case["H"], case["W"]
    m = model_from_tensors(state_tensors(st), H, W, dev(), swd_patcht_size=case["pt"])
    batched = {k: ([v] if isinstance(v, str) else torch.tensor([v])) for k, v in cfg.items()}

    def cuda_pass(wr, wa):
        m.args.rgb_smooth_loss_weight, m.args.a_smooth_loss_weight = wr, wa
        m.zero_grad()
        m.train()
        _, ex = m(H, W, ext.to(dev()), intr.to(dev()), res=res.to(dev()), losscfg=batched)
        loss = ex["swd"].mean()
        for k, w_ in (("rgb_smooth", wr), ("a_smooth", wa)):
            if w_ > 0:
                loss = loss + w_ * ex[k].mean()
        loss.backward()
        return ex, loss

    def oracle_pass(wr, wa, nn_override=None):
        a = st.atlas.double().requires_grad_(True)
        ad = st.atlas_dyn.double().requires_grad_(True)
        c2 = dict(cfg, nn_override=nn_override) if nn_override is not None else cfg
        extra, aux = MO.forward_train(st, H, W, ext, intr, res, c2, swd_patcht_size=case["pt"], atlas=a, atlas_dyn=ad,
                                      rgb_smooth=wr > 0, a_smooth=wa > 0)
        MO.total_loss(extra, wr, wa).backward()
        return extra, aux, a.grad, ad.grad

    # 1) without the (sign-valued) smoothness gradients: strict comparison
    ex, _ = cuda_pass(0.0, 0.0)
    extra, aux, ga, gad = oracle_pass(0.0, 0.0)
    nn_c = m.losses["gpnn_lm"].last_nn.cpu().long()
    ycrop = res.permute(0, 2, 1, 3, 4).double()[..., :aux["x"].shape[-2], :aux["x"].shape[-1]]
    mism, bad = LL.tie_margin_ok(aux["x"].detach(), ycrop, aux["nn"], nn_c, case["p"], case["pt"], case["s"], case["st"],
                                 case["alpha"])
    assert bad == 0, f"{bad} NN mismatches that are not fp64 near-ties"
    assert mism <= 2, f"{mism} near-tie NN mismatches out of {nn_c.numel()}"
    nn_over = nn_c if mism else None
    if mism:            # exact near-tie(s): continue the comparison with CUDA's (equally valid) matches
        extra, aux, ga, gad = oracle_pass(0.0, 0.0, nn_over)
    assert abs(float(ex["swd"]) - float(extra["swd"])) < RTOL * abs(float(extra["swd"]))
    assert relerr(m.atlas_dyn.grad.cpu(), gad) < 2e-4
    if m.mesh_pack().n_static:
        assert relerr(m.atlas.grad.cpu(), ga) < 2e-4
    # 2) with smoothness: d|a-b| = sign(a-b) flips wherever two neighbouring values agree to ~1e-6, so a
    #    handful of texels may differ by up to 2*w_smooth; everything else must agree as tightly as above
    ex, loss_c = cuda_pass(0.2, 0.2)
    extra, aux, ga, gad = oracle_pass(0.2, 0.2, nn_over)
    for k in ("swd", "rgb_smooth", "a_smooth"):
        assert abs(float(ex[k]) - float(extra[k])) < RTOL * abs(float(extra[k])), k
    T_ = st.atlas_dyn.shape[0]
    w_max = 2 * 0.2 * 3.5 / (T_ * (H - 1) * (W - 1) * case["D"])             # two sign flips of one pair term
    for got, ref in ((m.atlas_dyn.grad.cpu().double(), gad),) + (((m.atlas.grad.cpu().double(), ga),)
                                                                  if m.mesh_pack().n_static else ()):
        err = (got - ref).abs()
        assert float(err.max()) <= 4 * w_max + 2e-4 * float(ref.abs().max())
        n_off = int((err > 2e-4 * ref.abs().max()).sum())
        assert n_off <= max(64, 2e-4 * ref.numel()), f"{n_off} texel gradients off"
    # FusedLoopStep on a fresh copy of the model, separate kernels (the gradient buffer can be inspected)
    m2 = model_from_tensors(state_tensors(st), H, W, dev(), swd_patcht_size=case["pt"])
    step = FusedLoopStep(m2, fused="off")
    out = step.step(H, W, ext.to(dev()), intr.to(dev()), res.to(dev()), cfg, lr=0.01)
    assert abs(float(out["loss"]) - float(loss_c)) < 1e-5 * abs(float(loss_c))
    g_dyn = step._buf["g_dyn"].clone()                                  # (the buffer is reused by the next step)
    assert relerr(g_dyn.cpu(), m.atlas_dyn.grad.cpu()) < 1e-5           # same kernels, same inputs
    p_ref, _, _ = MO.adam_step(st.atlas_dyn, g_dyn.cpu(), torch.zeros_like(st.atlas_dyn), torch.zeros_like(st.atlas_dyn),
                               1, 0.01)
    assert float((m2.atlas_dyn.detach().cpu() - p_ref).abs().max()) < 1e-6
    # a second step keeps running (Adam state, buffers re-used)
    out2 = step.step(H, W, ext.to(dev()), intr.to(dev()), res.to(dev()), cfg, lr=0.01)
    assert torch.isfinite(out2["loss"]) and step.t == 2
    # the default path (backward + Adam in one persistent kernel): same losses, same first moment, same update
    m3 = model_from_tensors(state_tensors(st), H, W, dev(), swd_patcht_size=case["pt"])
    step3 = FusedLoopStep(m3)
    out3 = step3.step(H, W, ext.to(dev()), intr.to(dev()), res.to(dev()), cfg, lr=0.01)
    assert step3.last_schedule is not None
    assert abs(float(out3["loss"]) - float(out["loss"])) < 1e-6 * abs(float(out["loss"]))
    assert relerr(step3._state["atlas_dyn"][0].cpu(), 0.1 * g_dyn.cpu()) < 1e-5
    d = (m3.atlas_dyn.detach().cpu() - p_ref).abs()                  # Adam's eps = 6e-8: a ~0 gradient may flip sign
    assert float((d > 1e-5).float().mean()) < 2e-4 and float(d.median()) < 1e-7
    out3b = step3.step(H, W, ext.to(dev()), intr.to(dev()), res.to(dev()), cfg, lr=0.01)
    assert abs(float(out3b["loss"]) - float(out2["loss"])) < 2e-5 * abs(float(out2["loss"]))


def test_render_edge_cases():
    """Ragged sizes (not multiples of the 32x8 tile), a single pixel row, one frame, frame subsets."""
    from videoloop3d_b200.testing import model_from_tensors
    for (H, W, T, D) in [(1, 33, 1, 3), (9, 1, 2, 1), (31, 63, 3, 32), (8, 32, 5, 2)]:
        st = MO.sparse_state(H, W, D, 4, 5, T, 1.0, 10.0, tile=4, occupancy=0.8, dyn_frac=0.5,
                             h_scale=max(1.5, 6 / H), w_scale=max(1.5, 6 / W), seed=H + W)
        ext = torch.eye(4)[None]
        f = 0.9 * max(W, 8)
        intr = torch.tensor([[f, 0, W / 2 + 0.2], [0, f, H / 2 - 0.3], [0, 0, 1.]])[None]
        m = model_from_tensors(state_tensors(st), H, W, dev())
        m.eval()
        ts = list(range(T))[::-1]
        with torch.no_grad():
            rgb, _ = m(H, W, ext.to(dev()), intr.to(dev()), ts=ts)
        ref, _ = MO.render(st, H, W, ext, intr, ts)
        assert float((rgb.cpu().permute(0, 2, 3, 1) - ref).abs().max()) < 1e-4
        with pytest.raises(IndexError):
            m(H, W, ext.to(dev()), intr.to(dev()), ts=[T])
