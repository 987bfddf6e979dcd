"""CPU: pin the oracle restatement (oracle/*.py) to the golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  No GPU, no /root/reference needed."""
import numpy as np
import pytest
import torch

from oracle import looploss_oracle as LL
from oracle import mpv_oracle as MO
from util import cfg_from_golden, load_golden, relerr, state_from_golden


def test_pointwise_known_answers():
    g = load_golden("pointwise")
    alpha, content = torch.as_tensor(g["alpha"]), torch.as_tensor(g["content"])
    bw = torch.cat([alpha[..., :1], alpha[..., 1:] * torch.cumprod(1 - alpha, -1)[..., :-1]], -1)
    assert relerr(bw, g["bw"]) < 1e-6
    assert relerr((content * bw[..., None]).sum(-2), g["rgb"]) < 1e-6
    assert np.allclose(MO.make_depths(8, 1.0, 10.0).numpy(), g["depths"])
    r = torch.as_tensor(g["r"])
    for rou in ("mse", "abs", "0", "2", "-2", "1"):
        assert relerr(LL.robust_lossfun(r, rou, 0.1), g["rho_" + rou]) < 1e-6, rou


@pytest.mark.parametrize("name", ["render_dense", "render_sparse"])
def test_render_matches_reference(name):
    g = load_golden(name)
    st = state_from_golden(g)
    H, W, T = int(g["H"]), int(g["W"]), int(g["T"])
    rgb, var = MO.render(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]), range(T))
    assert var["K"] == int(g["K"])
    assert torch.equal(var["hitmask"].reshape(H, W, -1).sum(-1), torch.as_tensor(g["pix_to_face"][0] >= 0).sum(-1))
    assert float((rgb - torch.as_tensor(g["rgb"])).abs().max()) < 2e-5
    assert float((var["mpi"] - torch.as_tensor(g["mpi"])).abs().max()) < 2e-5
    assert float((var["alpha"] - torch.as_tensor(g["alpha"])).abs().max()) < 2e-5
    assert float((var["blend_weight"] - torch.as_tensor(g["blend_weight"])).abs().max()) < 2e-5
    rgb2, _ = MO.render(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]), [T - 1, 0])
    assert float((rgb2.permute(0, 3, 1, 2) - torch.as_tensor(g["rgb_eval_ts"])).abs().max()) < 2e-5


@pytest.mark.parametrize("name", ["step_dense_refcfg", "step_sparse_othercfg"])
def test_train_step_matches_reference(name):
    g = load_golden(name)
    st = state_from_golden(g)
    cfg = cfg_from_golden(g)
    H, W = int(g["H"]), int(g["W"])
    a = st.atlas.double().requires_grad_(True)
    ad = st.atlas_dyn.double().requires_grad_(True)
    extra, aux = MO.forward_train(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]),
                                  torch.as_tensor(g["res"]), cfg, atlas=a, atlas_dyn=ad)
    for k, v in extra.items():
        assert abs(float(v) - float(g["extra_" + k])) < 2e-6 * max(1.0, abs(float(v))), k
    loss = MO.total_loss(extra, float(g["rgb_smooth_w"]), float(g["a_smooth_w"]))
    assert abs(float(loss) - float(g["loss"])) < 5e-6
    loss.backward()
    assert relerr(ad.grad, g["grad_atlas_dyn"]) < 1e-4
    if g["grad_atlas"].size > 4:
        assert relerr(a.grad, g["grad_atlas"]) < 1e-4
    assert float((aux["y2x"] - torch.as_tensor(g["y2x"])).abs().max()) < 1e-6
    # Adam (eps=6e-8) on the reference's own gradients reproduces its parameter update bit for bit
    p0, gr = torch.as_tensor(g["atlas_dyn"]), torch.as_tensor(g["grad_atlas_dyn"])
    p1, _, _ = MO.adam_step(p0, gr, torch.zeros_like(p0), torch.zeros_like(p0), 1, float(g["lr"]))
    assert float((p1 - torch.as_tensor(g["new_atlas_dyn"])).abs().max()) < 1e-7


@pytest.mark.parametrize("name", ["step_dense_terms", "step_sparse_terms"])
def test_optional_terms_match_reference(name):
    """bg_color (MPV.py:455-461), sparsity (:511-515), density (:533-536), d_smooth / disp (:463-464, 538-551):
    off in the shipped stage-2 configs, pinned here against the unmodified reference run with them on."""
    g = load_golden(name)
    st = state_from_golden(g)
    cfg = cfg_from_golden(g)
    H, W = int(g["H"]), int(g["W"])
    a = st.atlas.double().requires_grad_(True)
    ad = st.atlas_dyn.double().requires_grad_(True)
    bg = MO.parse_bg_color(str(g["bg_color"])) if str(g["bg_color"]) else None
    extra, aux = MO.forward_train(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]),
                                  torch.as_tensor(g["res"]), cfg, atlas=a, atlas_dyn=ad, sparsity=True, density=True,
                                  d_smooth=True, bg_color=bg)
    assert set(extra) == {"swd", "rgb_smooth", "a_smooth", "sparsity", "density", "d_smooth"}
    for k, v in extra.items():
        assert abs(float(v) - float(g["extra_" + k].reshape(-1)[0])) < 2e-6 * max(1.0, abs(float(v))), k
    loss = MO.total_loss(extra, float(g["rgb_smooth_w"]), float(g["a_smooth_w"]), sparsity=float(g["sparsity_w"]),
                         density=float(g["density_w"]), d_smooth=float(g["d_smooth_w"]))
    assert abs(float(loss) - float(g["loss"])) < 5e-6
    loss.backward()
    assert relerr(ad.grad, g["grad_atlas_dyn"]) < 1e-4
    if g["grad_atlas"].size > 4:
        assert relerr(a.grad, g["grad_atlas"]) < 1e-4


@pytest.mark.parametrize("name", ["stage1_loopmask", "stage1_bg_normdepth"])
def test_stage1_render_and_forward_match_reference(name):
    """SURVEY §8(f) N4, second half: `MPI.MPMesh.render / forward` (MPI.py:452-652) of the unmodified reference —
    rgb + loop-mask label, normalised disparity, every extra term and the gradients of atlas / atlas_mask."""
    g = load_golden(name)
    st = state_from_golden(g)
    H, W = int(g["H"]), int(g["W"])
    a = st.atlas.double().requires_grad_(True)
    am = torch.as_tensor(g["atlas_mask"]).double().requires_grad_(True) if bool(g["loop_mask"]) else None
    bg = MO.parse_bg_color(str(g["bg_color"])) if str(g["bg_color"]) else None
    rgbl, extra, var = MO.forward_stage1(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]),
                                         float(g["near"]), float(g["far"]), edge_scale=float(g["edge_scale"]), atlas=a,
                                         atlas_mask=am, bg_color=bg,
                                         normalize_blendweight_fordepth=bool(g["normalize_blendweight_fordepth"]))
    assert tuple(rgbl.shape) == tuple(g["rgbl"].shape)
    assert relerr(rgbl.detach(), g["rgbl"]) < 2e-5
    assert relerr(var["disp_norm"].detach(), g["disp_norm"]) < 2e-5
    assert relerr(var["blend_weight"].detach(), g["blend_weight"]) < 2e-5
    assert relerr(var["alpha"].detach(), g["alpha"]) < 2e-5
    if am is not None:
        assert relerr(var["loopmask3d"].detach(), g["loopmask3d"]) < 2e-5
    assert set(extra) == {k[6:] for k in g if k.startswith("extra_")}
    loss = (rgbl * torch.as_tensor(g["g_up"])).mean()
    for k, v in extra.items():
        assert abs(float(v) - float(g["extra_" + k].reshape(-1)[0])) < 2e-6 * max(1.0, abs(float(v))), k
        loss = loss + v * float(g["w_" + k])
    assert abs(float(loss) - float(g["loss"])) < 2e-6
    loss.backward()
    assert relerr(a.grad, g["grad_atlas"]) < 1e-4
    if am is not None:
        assert relerr(am.grad, g["grad_atlas_mask"]) < 1e-4


@pytest.mark.parametrize("name", ["loss_lm_alpha0", "loss_lm_noalpha", "loss_direct_p7", "loss_lm_abs"])
@pytest.mark.parametrize("mode", ["exact64", "ref32"])
def test_loss_matches_reference(name, mode):
    g = load_golden(name)
    cfg = cfg_from_golden(g)
    x = torch.as_tensor(g["x"]).double().requires_grad_(True)
    y = torch.as_tensor(g["y"]).double()
    fn = LL.gpnn_lowmem if str(g["cls"]) == "Patch3DGPNNLowMemLoss" else LL.gpnn_direct
    loss, aux = fn(x, y, nn_mode=mode, **cfg)
    (gr,) = torch.autograd.grad(loss, x)
    assert torch.equal(aux["nn"], torch.as_tensor(g["nn"]))            # NN indices: bit-exact
    assert abs(float(loss) - float(g["loss"])) < 2e-6
    assert float((aux["y2x"] - torch.as_tensor(g["y2x"])).abs().max()) < 1e-6
    assert torch.equal(aux["weight"].float(), torch.as_tensor(g["weight"]))
    assert relerr(gr, g["grad_x"]) < 1e-5


def test_macro_block_loop_is_a_memory_trick():
    g = load_golden("loss_lm_alpha0")
    cfg = cfg_from_golden(g)
    x, y = torch.as_tensor(g["x"]).double(), torch.as_tensor(g["y"]).double()
    l0, a0 = LL.gpnn_lowmem(x, y, **cfg)
    l1, a1 = LL.gpnn_lowmem(x, y, use_macro_blocks=True, **cfg)
    assert abs(float(l0 - l1)) < 1e-12 and float((a0["y2x"] - a1["y2x"]).abs().max()) < 1e-12
    assert torch.equal(a0["weight"], a1["weight"])


def test_nnerr_matches_reference():
    """evaluations/NNMSE.compute_nnerr of the unmodified reference."""
    import warnings
    g = load_golden("nnerr")
    src, tar = torch.as_tensor(g["src"]).double(), torch.as_tensor(g["tar"]).double()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(3):
            p, s, pt, st = (int(v) for v in g[f"cfg{i}"])
            e = LL.compute_nnerr(src, tar, p, s, pt, st, macro_block=25)
            assert abs(e - float(g[f"err{i}"])) < 1e-6 * float(g[f"err{i}"])


@pytest.mark.parametrize("name", ["step_dense_refcfg", "step_sparse_othercfg"])
def test_reference_operator_sequence_matches_reference(name):
    """oracle/torch_ref_ops.py (the timed `bench.py --impl reference --ref-device cuda` baseline: grid_sample /
    masked_scatter / cumprod / unfold / bmm / index_add in the reference's order, fp32) reproduces the unmodified
    reference's losses and gradients of a whole stage-2 step."""
    from oracle import torch_ref_ops as RO
    g = load_golden(name)
    st = state_from_golden(g)
    cfg = cfg_from_golden(g)
    H, W = int(g["H"]), int(g["W"])
    tabs = RO.raster_tables(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]), "cpu")
    a = st.atlas.float().requires_grad_(True)
    ad = st.atlas_dyn.float().requires_grad_(True)
    total, extra, _ = RO.forward_train_ops(a, ad, tabs, H, W, torch.as_tensor(g["res"]), cfg, st.mpi_d,
                                           rgb_smooth_w=float(g["rgb_smooth_w"]), a_smooth_w=float(g["a_smooth_w"]))
    for k, v in extra.items():
        assert abs(float(v) - float(g["extra_" + k].reshape(-1)[0])) < 1e-5 * max(1.0, abs(float(v))), k
    assert abs(float(total) - float(g["loss"])) < 1e-5
    total.backward()
    assert relerr(ad.grad, g["grad_atlas_dyn"]) < 1e-3
