"""GPU: the SURVEY §8(f) "next" rows built on the hot-path kernels — N3 NN-error metric
(evaluations/NNMSE.py) against the unmodified reference's values, N1 uint8 epilogue (utils.py:17), and
the lod()/checkpoint plumbing of N2 (MPV.py:140-198, 290-304) as an API-level smoke test."""
import numpy as np
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu


def test_compute_nnerr_matches_reference_golden():
    from videoloop3d_b200 import compute_nnerr
    g = load_golden("nnerr")
    dev = torch.device("cuda:0")
    src, tar = torch.as_tensor(g["src"]).to(dev), torch.as_tensor(g["tar"]).to(dev)
    for i in range(3):
        p, s, pt, st = (int(v) for v in g[f"cfg{i}"])
        e = compute_nnerr(src, tar, p, s, pt, st, macro_block=25)
        assert abs(e - float(g[f"err{i}"])) < 1e-5 * float(g[f"err{i}"]), (i, e)


def test_to8b_matches_numpy():
    from videoloop3d_b200 import to8b
    x = torch.randn(3, 3, 37, 53, device="cuda:0") * 0.7 + 0.5
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 1.0 - 1e-7, 254.999 / 255], device="cuda:0")
    ref = (255 * np.clip(x.cpu().numpy(), 0, 1)).astype(np.uint8).transpose(0, 2, 3, 1)      # utils.py:17
    assert np.array_equal(to8b(x).cpu().numpy(), ref)


def test_lod_and_state_dict_roundtrip():
    from oracle import mpv_oracle as MO
    from test_gpu_parity import state_tensors
    from videoloop3d_b200.testing import model_from_tensors
    dev = torch.device("cuda:0")
    H, W = 32, 48
    st = MO.sparse_state(H, W, 4, 5, 7, 3, 1.0, 10.0, tile=8, occupancy=0.8, dyn_frac=0.5, h_scale=1.3, w_scale=1.3, seed=3)
    m = model_from_tensors(state_tensors(st), H, W, dev)
    m.atlas_grid_h, m.atlas_grid_w = st.atlas.shape[-2] // 8, st.atlas.shape[-1] // 8
    m.atlas_full_h, m.atlas_full_w = st.atlas.shape[-2:]
    m.atlas_grid_dyn_h, m.atlas_grid_dyn_w = st.atlas_dyn.shape[-2] // 8, st.atlas_dyn.shape[-1] // 8
    m.atlas_full_dyn_h, m.atlas_full_dyn_w = st.atlas_dyn.shape[-2:]
    ext = torch.eye(4, device=dev)[None]
    intr = torch.tensor([[40., 0, 24.2], [0, 40., 15.7], [0, 0, 1]], device=dev)[None]
    m.eval()
    with torch.no_grad():
        full, _ = m(H, W, ext, intr, ts=[0, 2])
    m.lod(0.5)                                                            # 8x8 tiles -> 4x4 tiles, uvs re-aligned
    assert m.atlas.shape[-1] == st.atlas.shape[-1] // 2 and m.atlas_dyn.shape[-2] == st.atlas_dyn.shape[-2] // 2
    assert m.atlas_dyn.is_contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        half, _ = m(H, W, ext, intr, ts=[0, 2])
    assert tuple(half.shape) == tuple(full.shape) and bool(torch.isfinite(half).all())
    assert float((half - full).abs().mean()) < 0.25                       # same scene at half the texel density
    sd = m.state_dict()
    for k in ("atlas", "atlas_dyn", "uvs", "uvs_dyn", "_verts", "faces", "faces_dyn", "uvfaces", "uvfaces_dyn", "planedepth",
              "ref_extrin", "ref_intrin", "self.is_sparse", "self.atlas_full_w", "self.atlas_grid_h", "self.has_dyn",
              "self.atlas_full_dyn_h", "self.atlas_grid_dyn_w"):
        assert k in sd, k                                                 # MPV.py:290-304
    opt = m.get_optimizer(step=0)
    assert len(opt.param_groups) == 2 and opt.param_groups[0]["eps"] == 6e-8
    assert abs(m.get_lrate(1000)[0][1] - m.args.lrate * 0.1 ** (1000 / (m.args.lrate_decay * 1000))) < 1e-12


def test_sharded_scale_invariant_matches_direct():
    """block-wise frame sums + vl3d_scale_invariant_presum == vl3d_scale_invariant (MPV.py:499-504)."""
    from videoloop3d_b200 import ops
    g = torch.Generator(device="cuda:0").manual_seed(3)
    T, F_, H, W = 5, 11, 19, 23
    rgb = torch.rand((T + 2, 3, H, W), device="cuda:0", generator=g)
    res = torch.rand((F_, 3, H, W), device="cuda:0", generator=g)
    direct = float(ops.scale_invariant(rgb, T, res))
    ref = float((torch.exp(torch.log((res.double().mean(0) + 0.01) / (rgb[:T].double().mean(0) + 0.01)).mean()) + 3) / 4)
    assert abs(direct - ref) < 1e-5 * abs(ref)
    for world in (2, 3):
        b = [(F_ * r) // world for r in range(world + 1)]
        total = sum(ops.frame_sum(res[b[r]:b[r + 1]].contiguous()) for r in range(world))
        got = float(ops.scale_invariant_presum(rgb, T, total.contiguous(), F_))
        assert abs(got - direct) < 2e-6 * abs(direct), (world, got, direct)


def test_checkpoint_load_lod_render_matches_reference():
    """N2 + N1 end to end through the C ABI: the unmodified reference loaded a stage-1 checkpoint with
    `init_from_mpi`, rendered frames through the eval branch of `forward`, applied `lod(0.5)` / `lod(1.0)` and
    rendered again, and loaded the static-only branch (oracle/make_golden.py::golden_ckpt); the same calls on
    the CUDA path give the same frames (1e-4 relative)."""
    from util import ckpt_dict, ckpt_model, relerr
    g = load_golden("ckpt")
    dev = torch.device("cuda:0")
    H, W, T = int(g["H"]), int(g["W"]), int(g["T"])
    ext, intr = torch.as_tensor(g["tar_extrin"]).to(dev), torch.as_tensor(g["tar_intrin"]).to(dev)
    m = ckpt_model(g, dev)
    m.init_from_mpi(ckpt_dict(g, "stage1_"))
    m.atlas_dyn.data = m.atlas_dyn.data + torch.as_tensor(g["noise"]).to(dev)
    m.eval()
    with torch.no_grad():
        rgb, _ = m(H, W, ext, intr, ts=[0, T - 1])
        assert relerr(rgb.cpu(), g["rgb_full"]) < 1e-4
        m.lod(0.5)
        rgb, _ = m(H, W, ext, intr, ts=[T - 1, 1])
        assert relerr(rgb.cpu(), g["rgb_half"]) < 1e-4
        m.lod(1.0)
        rgb, _ = m(H, W, ext, intr, ts=[1])
        assert relerr(rgb.cpu(), g["rgb_up"]) < 1e-4
        m2 = ckpt_model(g, dev)
        m2.init_from_mpi(m.state_dict())
        m2.eval()
        assert torch.equal(m2(H, W, ext, intr, ts=[1])[0], rgb)
        m3 = ckpt_model(g, dev)
        m3.init_from_mpi({k: v for k, v in ckpt_dict(g, "stage1_").items() if "dyn" not in k})
        m3.eval()
        rgb, _ = m3(H, W, ext, intr, ts=[0, T - 1])
        assert relerr(rgb.cpu(), g["rgb_static"]) < 1e-4


def test_run_iter_on_dataset_items_matches_reference_step():
    """S1 + S3 (train_3dvid.py:214-255 fed by MVVidPatchDataset items): `make_run_iter` driven by a batch in the
    DataLoader format reproduces the unmodified reference's step (golden `step_dense_refcfg`: total loss 1e-4,
    parameter update in the same direction), and `FusedLoopStep` fed from a GPU-resident dataset (crops are
    strided views of the resident video) computes the same loss for every item."""
    from util import cfg_from_golden
    from test_gpu_parity import model_from_golden
    from videoloop3d_b200 import FusedLoopStep, MVVidPatchDataset, make_run_iter
    g = load_golden("step_dense_refcfg")
    cfg = cfg_from_golden(g)
    dev = torch.device("cuda:0")
    kw = dict(rgb_smooth_loss_weight=float(g["rgb_smooth_w"]), a_smooth_loss_weight=float(g["a_smooth_w"]),
              swd_patcht_size=int(cfg["patcht_size"]), add_intrin_noise=False)
    m = model_from_golden(g, **kw)
    H, W = int(g["H"]), int(g["W"])
    ext = torch.as_tensor(g["tar_extrin"])
    pose = torch.linalg.inv(ext)[:, :3, :4]                               # run_iter inverts it again (utils.py:211-219)
    batched = {k: ([v] if isinstance(v, str) else torch.tensor([v])) for k, v in cfg.items()}
    item = (torch.tensor([0]), torch.tensor([0]), pose, torch.as_tensor(g["tar_intrin"]), torch.as_tensor(g["res"]), batched)
    opt = m.get_optimizer(step=0)
    for grp in opt.param_groups:
        grp["lr"] = float(g["lr"])
    p0 = m.atlas_dyn.data.clone()
    loss = make_run_iter(m.args, m, dev)(0, opt, item)
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    d_ours = (m.atlas_dyn.data - p0).cpu().contiguous()
    d_ref = torch.as_tensor(g["new_atlas_dyn"]) - torch.as_tensor(g["atlas_dyn"])
    big = torch.as_tensor(g["grad_atlas_dyn"]).abs() > 1e-6              # Adam(eps=6e-8) amplifies noise on ~zero gradients
    assert float((d_ours[big] - d_ref[big]).abs().max()) < 2e-3 * float(g["lr"])
    assert int(big.sum()) > 1000

    # GPU-resident dataset -> fused step, crops are views
    rng = np.random.default_rng(0)
    F_ = g["res"].shape[1]
    videos = [rng.integers(0, 256, (F_, 2 * H, 2 * W, 3), dtype=np.uint8)]
    poses = pose.clone()
    intr = torch.as_tensor(g["tar_intrin"]).clone()
    intr[0, 0, 2] += W / 2
    intr[0, 1, 2] += H / 2
    plain = {k: v for k, v in cfg.items()}
    ds = MVVidPatchDataset((2 * H, 2 * W), videos, (H, W), (H, W), poses, intr, loss_configs=[plain], device=dev)
    assert len(ds) == 4 and ds.videos[0].is_cuda
    m2 = model_from_golden(g, **kw)
    step = FusedLoopStep(m2)
    m3 = model_from_golden(g, **kw)
    m3.train()
    for b in ds.batches(shuffle=False):
        w0, h0, b_pose, b_intr, crops, bcfg = b
        assert crops.is_cuda and not crops.is_contiguous()
        from videoloop3d_b200 import pose2extrin_torch
        b_ext = pose2extrin_torch(b_pose)
        out = step.step(H, W, b_ext, b_intr, crops, plain, lr=0.0, optimise=False)
        _, extra = m3(H, W, b_ext.to(dev), b_intr.to(dev), res=crops, losscfg=bcfg)
        ref = extra["swd"].mean() + extra["rgb_smooth"].mean() * kw["rgb_smooth_loss_weight"] \
            + extra["a_smooth"].mean() * kw["a_smooth_loss_weight"]
        assert abs(float(out["loss"]) - float(ref)) < 1e-4 * abs(float(ref))


def test_u8_targets_convert_on_device_bit_exactly():
    """S3: `vid / 255` (train_3dvid.py:54) on the device (`vl3d_u8_to_unit`) gives the host conversion's bits for
    every byte value, for contiguous videos and for strided crops (aligned and unaligned), and the fused step fed
    with byte crops computes exactly what it computes from the float crops."""
    from videoloop3d_b200 import FusedLoopStep, ops
    from test_gpu_parity import model_from_golden
    from util import cfg_from_golden
    dev = torch.device("cuda:0")
    allv = torch.arange(256, dtype=torch.uint8).reshape(1, 1, 16, 16).expand(2, 3, 16, 16).contiguous()
    assert torch.equal(ops.u8_to_unit(allv.to(dev)).cpu(), allv / 255)
    g = torch.Generator().manual_seed(0)
    vid = torch.randint(0, 256, (5, 3, 37, 64), dtype=torch.uint8, generator=g)
    for crop in (vid, vid[..., 4:28, 8:48], vid[..., 3:30, 5:42], vid[1:4, :, :, 1:]):
        assert torch.equal(ops.u8_to_unit(_same_crop(vid.to(dev), vid, crop)).cpu(), crop / 255)
    gold = load_golden("step_dense_refcfg")
    cfg = cfg_from_golden(gold)
    H, W = int(gold["H"]), int(gold["W"])
    kw = dict(rgb_smooth_loss_weight=float(gold["rgb_smooth_w"]), a_smooth_loss_weight=float(gold["a_smooth_w"]),
              swd_patcht_size=int(cfg["patcht_size"]))
    F_ = gold["res"].shape[1]
    big = torch.randint(0, 256, (F_, 3, H + 6, W + 8), dtype=torch.uint8, generator=g).to(dev)
    crop_u8 = big[..., 2:2 + H, 4:4 + W]
    ext, intr = torch.as_tensor(gold["tar_extrin"]).to(dev), torch.as_tensor(gold["tar_intrin"]).to(dev)
    as_float = (crop_u8.cpu().float() / 255).to(dev)
    outs = []
    for res in (crop_u8[None], as_float[None]):
        step = FusedLoopStep(model_from_golden(gold, **kw))
        o = step.step(H, W, ext, intr, res, dict(cfg), lr=0.01)
        outs.append((float(o["loss"]), float(o["swd"]), step._buf["nn"].clone()))
        if res.dtype == torch.uint8:
            assert torch.equal(step._buf["res_f32"], as_float)
    # (the updated parameters are not compared bit for bit: the backward accumulates with float atomics)
    assert outs[0][:2] == outs[1][:2] and torch.equal(outs[0][2], outs[1][2])


def _same_crop(dev_vid, host_vid, host_crop):
    """The view of `dev_vid` that corresponds to `host_crop`, a basic-slicing view of `host_vid`."""
    off = host_crop.storage_offset() - host_vid.storage_offset()
    return dev_vid.as_strided(host_crop.shape, host_crop.stride(), dev_vid.storage_offset() + off)


def test_trivial_video_losses_match_their_formulas():
    """`MPMeshVid.losses['mse' | 'avg']` (utils_vid.py:437-445) through `vl3d_video_loss`: value and gradient against the
    reference's two-line torch expressions, equal and unequal frame counts."""
    import videoloop3d_b200 as V
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for tx, ty in ((5, 5), (7, 4), (3, 6)):
        x = torch.rand(1, 3, tx, 13, 17, generator=g).to(dev).requires_grad_(True)
        y = torch.rand(1, 3, ty, 13, 17, generator=g).to(dev)
        frm = min(tx, ty)
        for fn, ref in ((V.Patch3DMSE, lambda a: ((a[:, :, :frm] - y[:, :, :frm]) ** 2).mean()),
                        (V.Patch3DAvg, lambda a: ((a.mean(dim=2) - y.mean(dim=2)) ** 2).mean())):
            x.grad = None
            out = fn(x, y, patch_size=3)
            (out * 2.5).backward()
            got_g = x.grad.clone()
            x.grad = None
            want = ref(x)
            (want * 2.5).backward()
            assert abs(float(out) - float(want)) < 1e-6 * abs(float(want))
            assert float((got_g - x.grad).abs().max()) < 1e-6 * float(x.grad.abs().max())
