"""GPU: the optional per-ray terms of the composite — background blend (MPV.py:455-461), sparsity (MPV.py:511-515),
density (MPV.py:533-536), disparity / d_smooth (MPV.py:384-385,463-464,538-551) — against the golden vectors of the
unmodified reference run with those terms ON, and against the CPU oracle across views and model layouts.

Every shipped stage-2 config leaves them off; the CUDA path for them is csrc/terms.cu behind
`vl3d_composite_terms_fwd / _bwd` (an autograd node next to the main composite).  Tolerance: 1e-4 relative on values
(north_star), 5e-4 of the largest entry on gradients (the same bar as the main backward's golden test)."""
import pytest
import torch

from oracle import mpv_oracle as MO
from test_gpu_composite_sweep import MODELS, VIEWS, _build, _rot
from test_gpu_parity import model_from_golden, state_tensors
from util import cfg_from_golden, load_golden, relerr, state_from_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["step_dense_terms", "step_sparse_terms"])
def test_train_forward_backward_with_optional_terms_matches_reference_golden(name):
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    weights = {k: float(g[k + "_w"]) for k in ("rgb_smooth", "a_smooth", "sparsity", "density", "d_smooth")}
    m = model_from_golden(g, swd_patcht_size=int(cfg["patcht_size"]), bg_color=str(g["bg_color"]),
                          **{k + "_loss_weight": w for k, w in weights.items()})
    H, W = int(g["H"]), int(g["W"])
    ext = torch.as_tensor(g["tar_extrin"]).to(dev())
    intr = torch.as_tensor(g["tar_intrin"]).to(dev())
    res = torch.as_tensor(g["res"]).to(dev())
    batched = {k: ([v] if isinstance(v, str) else torch.tensor([v])) for k, v in cfg.items()}
    m.train()
    none, extra = m(H, W, ext, intr, res=res, losscfg=batched)
    assert none is None
    assert set(extra) == {"swd", "rgb_smooth", "a_smooth", "sparsity", "density", "d_smooth"}
    for k, v in extra.items():
        assert tuple(v.shape) == (1, 1)
        ref = float(g["extra_" + k].reshape(-1)[0])
        assert abs(float(v) - ref) < RTOL * abs(ref), (k, float(v), ref)
    lossobj = m.losses[cfg["loss_name"]]
    assert relerr(lossobj.last_y2x.cpu(), g["y2x"]) < 1e-5            # same NN matches as the reference
    loss = extra["swd"].mean()
    for k, w in weights.items():
        loss = loss + extra[k].mean() * w                              # train_3dvid.py:230-240
    assert abs(float(loss) - float(g["loss"])) < RTOL * abs(float(g["loss"]))
    loss.backward()
    assert relerr(m.atlas_dyn.grad.cpu(), g["grad_atlas_dyn"]) < 5e-4
    if g["grad_atlas"].size > 4:
        assert relerr(m.atlas.grad.cpu(), g["grad_atlas"]) < 5e-4
    # the fused step does not cover these terms and must say so instead of silently dropping them
    from videoloop3d_b200 import FusedLoopStep
    with pytest.raises(NotImplementedError):
        FusedLoopStep(m).step(H, W, ext.cpu(), intr.cpu(), res, cfg, 0.01)


def test_eval_render_blends_background_like_the_reference():
    g = load_golden("step_dense_terms")
    m = model_from_golden(g, bg_color=str(g["bg_color"]))
    st = state_from_golden(g)
    H, W = int(g["H"]), int(g["W"])
    T = st.atlas_dyn.shape[0]
    ext, intr = torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"])
    rgb_o, var_o = MO.render(st, H, W, ext, intr, range(T), bg_color=MO.parse_bg_color(str(g["bg_color"])))
    m.eval()
    with torch.no_grad():
        rgb_c, extra = m(H, W, ext.to(dev()), intr.to(dev()))
        rgb_r, var_r = m.render(H, W, (ext @ torch.inverse(st.ref_extrin)[None]).to(dev()), intr.to(dev()), list(range(T)))
    assert extra == {}
    assert float((rgb_c.cpu().double() - rgb_o.permute(0, 3, 1, 2)).abs().max()) < RTOL
    assert float((rgb_r.cpu().double() - rgb_o).abs().max()) < RTOL
    assert float((var_r["alpha"].cpu().double() - var_o["alpha"]).abs().max()) < RTOL


def test_random_background_follows_the_reference_rng_stream():
    """`bg_color='random'`: the background is `torch.rand(3)` from the global CPU generator, drawn once per render exactly as
    MPV.py:456-457 does, so a seeded run blends the same colours as the reference (and a second render a different one)."""
    g = load_golden("render_bg_random")
    m = model_from_golden(g, bg_color="random")
    H, W = int(g["H"]), int(g["W"])
    ext, intr = torch.as_tensor(g["tar_extrin"]).to(dev()), torch.as_tensor(g["tar_intrin"]).to(dev())
    m.eval()
    torch.manual_seed(int(g["seed"]))
    with torch.no_grad():
        rgb, _ = m(H, W, ext, intr)
        rgb2, _ = m(H, W, ext, intr)
    assert relerr(rgb.cpu(), g["rgb"]) < RTOL and relerr(rgb2.cpu(), g["rgb2"]) < RTOL
    assert float((rgb - rgb2).abs().max()) > 1e-3


@pytest.mark.parametrize("vname", sorted(VIEWS))
@pytest.mark.parametrize("mname", sorted(MODELS))
def test_terms_and_their_backward_match_oracle(mname, vname):
    """alpha / disp / sparsity sum and the gradients of a random functional of them, per view and layout (frame subsets
    with repeats, static-only, dynamic-only, a single plane, image sizes off the tile grid)."""
    from videoloop3d_b200 import ops
    from videoloop3d_b200.testing import model_from_tensors
    H, W = 37, 70
    seed = 3 + 17 * sorted(MODELS).index(mname) + 5 * sorted(VIEWS).index(vname)
    st = _build(mname, H, W, seed)
    v = VIEWS[vname]
    ext = torch.eye(4)
    ext[:3, :3] = _rot(*v["rot"])
    ext[:3, 3] = torch.tensor(v["trans"])
    f = 0.8 * W * v["fmul"]
    intr = torch.tensor([[f, 0, W / 2 + 0.37], [0, f, H / 2 - 0.21], [0, 0, 1.]])
    T = st.atlas_dyn.shape[0]
    ts = list(range(T))[::-1] + [0]
    D = st.mpi_d
    # oracle
    a = st.atlas.double().requires_grad_(True)
    ad = st.atlas_dyn.double().requires_grad_(True)
    _, var_o = MO.render(st, H, W, ext[None], intr[None], ts, atlas=a, atlas_dyn=ad)
    al = var_o["mpi"][..., -1]
    sp_o = (al.norm(dim=-1, p=1) / al.norm(dim=-1, p=2).clamp_min(1e-4)).sum() if var_o["K"] else torch.zeros((), dtype=torch.float64)
    gen = torch.Generator().manual_seed(seed)
    ga = torch.rand(var_o["alpha"].shape, generator=gen, dtype=torch.float64) - 0.3
    gd = torch.rand(var_o["alpha"].shape, generator=gen, dtype=torch.float64) - 0.5
    wsp = 0.37
    fun_o = (var_o["alpha"] * ga).sum() + (var_o["disp_norm"] * gd).sum() + wsp * sp_o
    fun_o.backward()
    # CUDA
    m = model_from_tensors(state_tensors(st), H, W, dev())
    extrin = (ext[None] @ torch.inverse(st.ref_extrin)[None])
    view = m.make_view(H, W, extrin, intr[None])
    atlas_dyn, atlas = m._texels()
    ts_t = m._ts_tensor(ts)
    alpha, disp, sp = m._render_terms(view, ts_t, len(ts), H, W, extrin, intr[None], want_disp=True, want_sparsity=True)
    assert float((alpha.detach().cpu().double() - var_o["alpha"].detach()).abs().max()) < RTOL
    scale_d = max(float(var_o["disp_norm"].abs().max()), 1e-6)
    assert float((disp.detach().cpu().double() - var_o["disp_norm"].detach()).abs().max()) < RTOL * scale_d
    assert abs(float(sp) - float(sp_o)) < RTOL * max(abs(float(sp_o)), 1.0)
    fun_c = (alpha * ga.to(dev()).float()).sum() + (disp * gd.to(dev()).float()).sum() + wsp * sp.sum()
    fun_c.backward()
    for name, got, ref in (("atlas_dyn", m.atlas_dyn.grad, ad.grad), ("atlas", m.atlas.grad, a.grad)):
        if ref is None or float(ref.abs().max()) == 0.0:
            assert got is None or float(got.abs().max()) < 1e-7, name
            continue
        got = got.cpu().double()
        assert float(got[:, :3].abs().max()) == 0.0, name                # these terms never touch the colour channels
        err = float((got - ref).abs().max())
        assert err < 3e-4 * float(ref.abs().max()), (name, err)


def test_terms_reject_bad_arguments():
    """C-ABI argument checks of the new entry points (error codes, no launch)."""
    from videoloop3d_b200 import _lib, ops
    g = load_golden("step_dense_terms")
    m = model_from_golden(g)
    H, W = int(g["H"]), int(g["W"])
    view = m.make_view(H, W, torch.eye(4)[None], torch.as_tensor(g["tar_intrin"]))
    atlas_dyn, atlas = m._texels()
    T = atlas_dyn.shape[0]
    with pytest.raises(_lib.Vl3dError, match="inverse-depth"):
        ops.composite_terms_fwd(view, m._pack, atlas_dyn, atlas, None, T, inv_depth=None, want_disp=True)
    with pytest.raises(_lib.Vl3dError, match="no output"):
        ops.composite_terms_fwd(view, m._pack, atlas_dyn, atlas, None, T, want_alpha=False)
    with pytest.raises(_lib.Vl3dError, match="sparsity_eps"):
        ops.composite_terms_fwd(view, m._pack, atlas_dyn, atlas, None, T, sparsity_eps=0.0)
    with pytest.raises(_lib.Vl3dError, match="bad T"):
        ops.composite_terms_fwd(view, m._pack, atlas_dyn, atlas, None, 0)
