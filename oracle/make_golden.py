"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run HERE (needs /root/reference; the GPU box only sees the committed .npz files):

    python -m oracle.make_golden

The reference's `MPV.MPMeshVid`, `utils_vid.Patch3DGPNN*Loss`, `utils_mpi.overcompose` etc. are
imported verbatim behind `oracle/ref_shims` (see its README for the "parity unpinned" caveat on
the pytorch3d rasteriser).  Inputs are seeded and stored together with the outputs, so tests can
replay them through the oracle restatement (CPU) and through the CUDA path (GPU).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_env  # noqa: E402
from oracle import mpv_oracle as MO  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(d):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def _view(seed, H, W, rot=0.05, trans=(0.08, -0.03, 0.02)):
    g = torch.Generator().manual_seed(seed)
    c, s = np.cos(rot), np.sin(rot)
    ext = torch.eye(4)
    ext[:3, :3] = torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float32)
    ext[:3, 3] = torch.tensor(trans)
    f = 0.8 * W
    intr = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1.]])
    intr[:2, 2] += torch.rand(2, generator=g) - 0.5          # train_3dvid.py:222-225
    return ext[None], intr[None]


def _load_state_into_reference(m, st):
    """Overwrite a reference MPMeshVid's tensors with an MPVState (same names, MPV.py:95-104)."""
    import torch.nn as nn
    m._verts.data = st.verts.clone()
    m.planedepth.data = st.planedepth.clone()
    m.ref_extrin.data = st.ref_extrin.clone()
    m.ref_intrin.data = st.ref_intrin.clone()
    m.uvs.data = st.uvs.clone()
    m.uvs_dyn.data = st.uvs_dyn.clone()
    m.uvfaces = st.uvfaces.clone()
    m.uvfaces_dyn = st.uvfaces_dyn.clone()
    m.faces = st.faces.clone()
    m.faces_dyn = st.faces_dyn.clone()
    m.register_parameter("atlas", nn.Parameter(st.atlas.clone()))
    m.register_parameter("atlas_dyn", nn.Parameter(st.atlas_dyn.clone()))
    m.frm_num = st.atlas_dyn.shape[0]
    m.is_sparse = True
    m.has_dyn = True


def _state_arrays(st):
    return dict(verts=st.verts, planedepth=st.planedepth, faces=st.faces, faces_dyn=st.faces_dyn, uvs=st.uvs,
                uvs_dyn=st.uvs_dyn, uvfaces=st.uvfaces, uvfaces_dyn=st.uvfaces_dyn, atlas=st.atlas,
                atlas_dyn=st.atlas_dyn, ref_extrin=st.ref_extrin, ref_intrin=st.ref_intrin,
                mpi_d=st.mpi_d, hv=st.hv, wv=st.wv)


def make_model(kind, H, W, D, hv, wv, T, seed, **kw):
    import MPV  # reference
    args = ref_env.make_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=kw.get("grid_h", 2),
                             mpv_frm_num=T, mpi_h_scale=kw.get("scale", 1.2), mpi_w_scale=kw.get("scale", 1.2),
                             add_intrin_noise=False, **kw.get("args", {}))
    f = 0.8 * W
    ref_intrin = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(seed)
    m = MPV.MPMeshVid(args, H, W, np.eye(4, dtype=np.float32), ref_intrin, 1.0, 10.0)
    if kind == "dense":
        st = MO.dense_state(H, W, D, hv, wv, kw.get("grid_h", 2), T, 1.0, 10.0, kw.get("scale", 1.2),
                            kw.get("scale", 1.2), seed=seed)
        # the oracle's constructor must reproduce the reference's own geometry exactly
        assert torch.allclose(st.verts, m._verts.data) and torch.equal(st.faces_dyn, m.faces_dyn)
        assert torch.allclose(st.uvs_dyn, m.uvs_dyn.data, atol=1e-7) and torch.allclose(st.planedepth, m.planedepth)
        st.atlas = st.atlas[:, :, :1, :1].clone()             # like init_from_mpi's dummy static (MPV.py:266)
    else:
        st = MO.sparse_state(H, W, D, hv, wv, T, 1.0, 10.0, tile=kw.get("tile", 6), occupancy=kw.get("occ", 0.6),
                             dyn_frac=0.5, h_scale=kw.get("scale", 1.2), w_scale=kw.get("scale", 1.2), seed=seed)
    _load_state_into_reference(m, st)
    return m, st, args


def golden_render(name, kind, H=24, W=40, D=4, hv=5, wv=7, T=3, seed=0, **kw):
    m, st, args = make_model(kind, H, W, D, hv, wv, T, seed, **kw)
    ext, intr = _view(seed, H, W)
    m.eval()
    ts = list(range(T))
    with torch.no_grad():
        extr = ext @ m.ref_extrin[None].inverse()
        rgb, var = m.render(H, W, extr, intr, ts)
        rgb_eval, _ = m(H, W, ext, intr, ts=[T - 1, 0])
    out = dict(H=H, W=W, T=T, tar_extrin=ext, tar_intrin=intr, rgb=rgb, mpi=var["mpi"], alpha=var["alpha"],
               blend_weight=var["blend_weight"], pix_to_face=var["pix_to_face"], K=var["mpi"].shape[-2],
               rgb_eval_ts=rgb_eval, **_state_arrays(st))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "K =", var["mpi"].shape[-2], "rgb mean", float(rgb.mean()))


def golden_render_bg_random(name, H=24, W=40, D=4, hv=5, wv=7, T=2, seed=16):
    """`bg_color='random'` (MPV.py:456-457 draws `torch.rand(3)` from the global CPU generator on every render): the
    eval forward of the unmodified reference under `torch.manual_seed(77)`."""
    m, st, args = make_model("dense", H, W, D, hv, wv, T, seed, args=dict(bg_color="random"))
    ext, intr = _view(seed, H, W)
    m.eval()
    torch.manual_seed(77)
    with torch.no_grad():
        rgb, _ = m(H, W, ext, intr)
        rgb2, _ = m(H, W, ext, intr)                               # the generator advances: a second, different background
    out = dict(H=H, W=W, T=T, tar_extrin=ext, tar_intrin=intr, rgb=rgb, rgb2=rgb2, seed=77, **_state_arrays(st))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "rgb mean", float(rgb.mean()), float(rgb2.mean()))


LOSS_CFG_REF = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=5, patcht_size=3, stride=2, stridet=1,
                    alpha=0.0, rou="-2", scaling=0.1, dist_fn="mse", macro_block=15, factor=1)
LOSS_CFG_OTHER = dict(loss_name="gpnn_lm", patch_size=3, patcht_size=3, stride=2, stridet=1,
                      alpha=10000.0, rou="-2", scaling=0.1, dist_fn="mse", macro_block=15, factor=1)


def _batched(cfg):
    """What the DataLoader hands to forward (MPV.py:494 un-batches with v[0])."""
    return {k: ([v] if isinstance(v, str) else torch.tensor([v])) for k, v in cfg.items()}


def golden_step(name, kind, cfg, H=24, W=40, D=4, hv=5, wv=7, T=6, F=9, seed=1, lr=0.05, **kw):
    m, st, args = make_model(kind, H, W, D, hv, wv, T, seed, **kw)
    ext, intr = _view(seed, H, W)
    g = torch.Generator().manual_seed(seed + 7)
    res = torch.rand(1, F, 3, H, W, generator=g)
    res = (res + res.roll(1, 1) + res.roll(2, 1)) / 3          # temporally smooth-ish
    m.train()
    params = [m.atlas, m.atlas_dyn]
    opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), eps=6e-8)   # MPV.py:213
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, extra = m(H, W, ext, intr, res=res, losscfg=_batched(cfg))
    extra_v = {k: v.detach().clone() for k, v in extra.items()}
    swd = extra.pop("swd").mean()                               # train_3dvid.py:230-244
    loss = swd
    for k, v in extra.items():
        w = getattr(args, f"{k}_loss_weight")
        if w > 0:
            loss = loss + v.mean() * w
    opt.zero_grad()
    loss.backward()
    g_atlas = m.atlas.grad.clone() if m.atlas.grad is not None else torch.zeros_like(m.atlas)
    g_dyn = m.atlas_dyn.grad.clone()
    opt.step()
    lossobj = m.losses[cfg["loss_name"]]
    out = dict(H=H, W=W, T=T, F=F, tar_extrin=ext, tar_intrin=intr, res=res, lr=lr, loss=loss.detach(),
               grad_atlas=g_atlas, grad_atlas_dyn=g_dyn, new_atlas=m.atlas.data, new_atlas_dyn=m.atlas_dyn.data,
               y2x=lossobj.last_y2x, weight=lossobj.last_weight,
               rgb_smooth_w=args.rgb_smooth_loss_weight, a_smooth_w=args.a_smooth_loss_weight,
               sparsity_w=args.sparsity_loss_weight, density_w=args.density_loss_weight,
               d_smooth_w=args.d_smooth_loss_weight, bg_color=str(args.bg_color),
               **{"extra_" + k: v for k, v in extra_v.items()},
               **{"cfg_" + k: v for k, v in cfg.items()}, **_state_arrays(st))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, {k: float(v) for k, v in extra_v.items()}, "loss", float(loss))


def golden_loss(name, cls, t=8, F=12, h=23, w=31, seed=3, smooth=False, **cfg):
    import utils_vid  # reference
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(1, 3, t, h, w, generator=g)
    y = torch.rand(1, 3, F, h, w, generator=g)
    if smooth:
        y = (y + y.roll(1, 2) + y.roll(2, 2)) / 3
        x = (y[:, :, :t] * 0.8 + 0.2 * x).contiguous()
    x.requires_grad_(True)
    lossobj = getattr(utils_vid, cls)()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = lossobj(x, y, **cfg)
    loss.backward()
    # NN indices of the global problem, straight from the reference's search (utils_vid.py:122-142)
    p, pt, s, st = cfg["patch_size"], cfg["patcht_size"], cfg["stride"], cfg["stridet"]
    hh, ww, tt = lossobj.last_y2x.shape[-2], lossobj.last_y2x.shape[-1], lossobj.last_y2x.shape[-3]
    with torch.no_grad():
        px = utils_vid.extract_3Dpatches(x[..., :tt, :hh, :ww], p, pt, s, st)
        b, c, d, ho, wo = px.shape
        px = px.permute(0, 3, 4, 2, 1).reshape(ho * wo, -1, 3, pt, p, p)
        py = utils_vid.extract_3Dpatches(y[..., :hh, :ww], p, pt, s, st)
        py = py.permute(0, 3, 4, 2, 1).reshape(ho * wo, -1, 3, pt, p, p)
        alpha = cfg.get("alpha", 1e10)
        nn = utils_vid.get_NN_indices_low_memory(px, py, None if alpha > 100 else alpha, 1024)
    out = dict(x=x.detach(), y=y, loss=loss.detach(), y2x=lossobj.last_y2x, weight=lossobj.last_weight,
               grad_x=x.grad, nn=nn.reshape(ho, wo, -1), cls=cls, **{"cfg_" + k: v for k, v in cfg.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "loss", float(loss))


def golden_pointwise(name):
    """overcompose (utils_mpi.py:92-107) + robust_lossfun (utils_vid.py:10-26) known answers."""
    import utils_mpi
    import utils_vid
    g = torch.Generator().manual_seed(5)
    alpha = torch.rand(2, 3, 4, 6, generator=g)
    content = torch.rand(2, 3, 4, 6, 3, generator=g)
    rgb, bw = utils_mpi.overcompose(alpha, content)
    r = torch.randn(257, generator=g) * 0.3
    out = dict(alpha=alpha, content=content, rgb=rgb, bw=bw, r=r, depths=utils_mpi.make_depths(8, 1.0, 10.0))
    for rou in ("mse", "abs", "0", "2", "-2", "1"):
        out["rho_" + rou] = utils_vid.robust_lossfun(r, rou, 0.1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name)


def golden_nnerr(name):
    """evaluations/NNMSE.compute_nnerr of the unmodified reference."""
    sys.path.insert(0, os.path.join(ref_env.REFERENCE_DIR, "evaluations"))
    import NNMSE  # reference
    g = torch.Generator().manual_seed(21)
    tar = torch.rand(1, 3, 14, 40, 52, generator=g)
    tar = (tar + tar.roll(1, 2) + tar.roll(1, 4)) / 3
    src = (tar[:, :, 3:12] * 0.7 + 0.3 * torch.rand(1, 3, 9, 40, 52, generator=g)).contiguous()
    out = dict(src=src, tar=tar)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, (p, s, pt, st) in enumerate([(5, 2, 7, 1), (11, 4, 5, 1), (17, 6, 3, 1)]):   # script_evaluate_ours.py:201-204
            out[f"err{i}"] = NNMSE.compute_nnerr(src, tar, p, s, pt, st, macro_block=25)
            out[f"cfg{i}"] = np.array([p, s, pt, st])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, [float(out[f"err{i}"]) for i in range(3)])


def golden_lod(name, H=32, W=48, D=4, hv=5, wv=7, T=3, tile=8, seed=3):
    """the unmodified reference's `MPMeshVid.lod` (MPV.py:140-198) on a tile-culled model —
    full -> 0.5 (down-sampling) -> 1.0 (up-sampling from the half-size tiles, as the pyramid schedule does,
    train_3dvid.py:259-272) — and on the dense layout.
    torchvision boundary: the reference pins torch==1.10 (torchvision 0.11), where `Resize(size)(tensor)` never
    anti-aliases; the torchvision installed here defaults to antialias=True for tensors.  For the DOWN-sampling
    call only, `Resize` is therefore constructed with the pinned version's behaviour (antialias=False) through a
    default-argument shim around the reference's call; the as-installed (anti-aliased) result is stored next to
    it (`*_aa`) to document the difference.  Up-sampling is identical under both."""
    import torchvision
    real_resize = torchvision.transforms.Resize

    class PinnedResize(real_resize):
        def __init__(self, size, *a, **k):
            k.setdefault("antialias", False)
            super().__init__(size, *a, **k)

    def grids(m, st):
        m.atlas_grid_h, m.atlas_grid_w = st.atlas.shape[-2] // tile, st.atlas.shape[-1] // tile
        m.atlas_full_h, m.atlas_full_w = st.atlas.shape[-2:]
        m.atlas_grid_dyn_h, m.atlas_grid_dyn_w = st.atlas_dyn.shape[-2] // tile, st.atlas_dyn.shape[-1] // tile
        m.atlas_full_dyn_h, m.atlas_full_dyn_w = st.atlas_dyn.shape[-2:]

    out = {}
    for tag, resize in (("", PinnedResize), ("_aa", real_resize)):
        m, st, _ = make_model("sparse", H, W, D, hv, wv, T, seed, tile=tile, occ=0.8, scale=1.3)
        grids(m, st)
        torchvision.transforms.Resize = resize
        try:
            m.lod(0.5)
            half = dict(atlas=m.atlas.data.clone(), atlas_dyn=m.atlas_dyn.data.clone(), uvs=m.uvs.data.clone(),
                        uvs_dyn=m.uvs_dyn.data.clone())
            m.lod(1.0)
        finally:
            torchvision.transforms.Resize = real_resize
        out.update({f"half{tag}_{k}": v for k, v in half.items()})
        out.update({f"full{tag}_atlas": m.atlas.data, f"full{tag}_atlas_dyn": m.atlas_dyn.data, f"full{tag}_uvs": m.uvs.data,
                    f"full{tag}_uvs_dyn": m.uvs_dyn.data})
        if tag == "":
            out.update(_state_arrays(st))
    # dense layout: 0.5 of the full atlas (pinned Resize), then back up to 1.0
    md, std, _ = make_model("dense", H, W, D, hv, wv, T, seed + 1, grid_h=2, scale=1.2)
    md.is_sparse = False
    torchvision.transforms.Resize = PinnedResize
    try:
        md.lod(0.5)
        out["dense_half_atlas_dyn"] = md.atlas_dyn.data.clone()
        md.lod(1.0)
        out["dense_full_atlas_dyn"] = md.atlas_dyn.data.clone()
    finally:
        torchvision.transforms.Resize = real_resize
    out["dense_atlas_dyn"] = std.atlas_dyn
    out.update(H=H, W=W, tile=tile, dense_grid_h=2, dense_scale=1.2, dense_seed=seed + 1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "half tiles", tuple(out["half_atlas_dyn"].shape), "aa differs by",
          float((out["half_atlas_dyn"] - out["half_aa_atlas_dyn"]).abs().max()),
          "| up-sampled", tuple(out["full_atlas_dyn"].shape))


def golden_ckpt(name, H=32, W=48, D=4, hv=5, wv=7, T=3, tile=8, seed=5):
    """SURVEY §8(f) N2, checkpoint format.  A stage-1 state dict as `MPI.MPMesh.state_dict` writes it after
    `sparsify_faces` (static tiles + ONE frame of dynamic tiles, MPI.py:207-221, 425-439) is loaded by the
    unmodified reference's `MPMeshVid.init_from_mpi` ("self.has_dyn" branch, MPV.py:241-262: the dynamic frame is
    replicated over T), run through the pyramid's `lod(0.5)` and `lod(1.0)` (train_3dvid.py:263-264), rendered
    through the eval branch of `forward` after each, saved with `state_dict()` and loaded again.  The static-only
    branch (MPV.py:264-288: the static MPI becomes the dynamic part) is rendered as well.  `lod(0.5)` runs under
    the same pinned-torchvision `Resize` default as golden_lod."""
    import MPV  # reference
    import torchvision
    real_resize = torchvision.transforms.Resize

    class PinnedResize(real_resize):
        def __init__(self, size, *a, **k):
            k.setdefault("antialias", False)
            super().__init__(size, *a, **k)

    st = MO.sparse_state(H, W, D, hv, wv, 1, 1.0, 10.0, tile=tile, occupancy=0.8, dyn_frac=0.5, h_scale=1.3,
                         w_scale=1.3, seed=seed)
    g = torch.Generator().manual_seed(seed)
    grid = lambda a: (int(a.shape[-2] // tile), int(a.shape[-1] // tile))
    stage1 = {"_verts": st.verts, "ref_extrin": st.ref_extrin, "ref_intrin": st.ref_intrin, "planedepth": st.planedepth,
              "uvs": st.uvs, "atlas": st.atlas, "uvfaces": st.uvfaces, "faces": st.faces, "self.is_sparse": True,
              "self.atlas_full_w": int(st.atlas.shape[-1]), "self.atlas_full_h": int(st.atlas.shape[-2]),
              "self.atlas_grid_h": grid(st.atlas)[0], "self.atlas_grid_w": grid(st.atlas)[1],
              "self.has_dyn": True, "uvs_dyn": st.uvs_dyn, "atlas_dyn": st.atlas_dyn, "uvfaces_dyn": st.uvfaces_dyn,
              "faces_dyn": st.faces_dyn, "self.atlas_full_dyn_w": int(st.atlas_dyn.shape[-1]),
              "self.atlas_full_dyn_h": int(st.atlas_dyn.shape[-2]), "self.atlas_grid_dyn_h": grid(st.atlas_dyn)[0],
              "self.atlas_grid_dyn_w": grid(st.atlas_dyn)[1]}
    args = ref_env.make_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpv_frm_num=T, mpi_h_scale=1.3,
                             mpi_w_scale=1.3, add_intrin_noise=False)
    f = 0.8 * W
    new_model = lambda: MPV.MPMeshVid(args, H, W, np.eye(4, dtype=np.float32),
                                      np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32), 1.0, 10.0)
    copy = lambda d: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in d.items()}
    torch.manual_seed(seed)
    m = new_model()
    m.init_from_mpi(copy(stage1))
    noise = 0.3 * torch.randn(m.atlas_dyn.shape, generator=g)
    m.atlas_dyn.data = m.atlas_dyn.data.clone() + noise           # un-alias the T expanded frames, make them differ
    ext, intr = _view(seed, H, W)
    m.eval()
    with torch.no_grad():
        rgb_full, _ = m(H, W, ext, intr, ts=[0, T - 1])
    torchvision.transforms.Resize = PinnedResize
    try:
        m.lod(0.5)
    finally:
        torchvision.transforms.Resize = real_resize
    with torch.no_grad():
        rgb_half, _ = m(H, W, ext, intr, ts=[T - 1, 1])
    m.lod(1.0)
    with torch.no_grad():
        rgb_up, _ = m(H, W, ext, intr, ts=[1])
    sd = m.state_dict()
    m2 = new_model()
    m2.init_from_mpi(copy(sd))
    m2.eval()
    with torch.no_grad():
        rgb_again, _ = m2(H, W, ext, intr, ts=[1])
    assert torch.equal(rgb_again, rgb_up)
    # static-only checkpoint (no "self.has_dyn"): the static MPI is loaded as the dynamic part
    static_only = {k: v for k, v in stage1.items() if "dyn" not in k}
    m3 = new_model()
    m3.init_from_mpi(copy(static_only))
    m3.eval()
    with torch.no_grad():
        rgb_static, _ = m3(H, W, ext, intr, ts=[0, T - 1])
    sd3 = m3.state_dict()
    key = lambda k: k.replace("self.", "self_")
    out = {"stage1_" + key(k): v for k, v in stage1.items()}
    out.update({"sd_" + key(k): v for k, v in sd.items()})
    out.update({"sd3_" + key(k): v for k, v in sd3.items()})
    out.update(H=H, W=W, T=T, D=D, hv=hv, wv=wv, tar_extrin=ext, tar_intrin=intr, noise=noise, rgb_full=rgb_full,
               rgb_half=rgb_half, rgb_up=rgb_up, rgb_static=rgb_static)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "keys", sorted(sd.keys()), "rgb", tuple(rgb_half.shape), float(rgb_half.mean()), float(rgb_static.mean()))


def golden_dataset(name, V=2, F=5, h_raw=20, w_raw=30, seed=7):
    """the unmodified reference's `MVVidPatchDataset` (train_3dvid.py:22-66) and
    `generate_patchinfo` (utils.py:115-134): every item of a small two-view capture, (a) resized + padded
    (16x24 target, 8x12 patches on a 6x8 stride grid) and (b) the small-frame branch (frame smaller than a patch)."""
    import train_3dvid as T3  # reference
    rng = np.random.default_rng(seed)
    videos = [rng.integers(0, 256, (F, h_raw, w_raw, 3), dtype=np.uint8) for _ in range(V)]
    poses = torch.tensor(rng.normal(size=(V, 3, 4)).astype(np.float32))
    intr = torch.tensor([[[25., 0, 15.], [0, 25., 10.], [0, 0, 1.]]]).repeat(V, 1, 1)
    intr[1, 0, 2] += 0.7
    cfgs = [dict(LOSS_CFG_REF), dict(LOSS_CFG_OTHER)]
    out = dict(V=V, poses=poses, intrins=intr, **{f"video{i}": v for i, v in enumerate(videos)})
    for tag, hw, psz, pst in (("a", (16, 24), (8, 12), (6, 8)), ("b", (10, 15), (16, 24), (8, 8)), ("c", (17, 26), (8, 12), (6, 8))):
        ds = T3.MVVidPatchDataset(hw, videos, psz, pst, poses, intr, loss_configs=cfgs)
        out.update({f"{tag}_hw": np.array(hw), f"{tag}_patch_size": np.array(psz), f"{tag}_patch_stride": np.array(pst),
                    f"{tag}_len": len(ds)})
        for i in range(len(ds)):
            w0, h0, pose, k, crops, cfg = ds[i]
            out.update({f"{tag}_{i}_wh": np.array([int(w0), int(h0)]), f"{tag}_{i}_pose": pose, f"{tag}_{i}_intrin": k,
                        f"{tag}_{i}_crops": crops, f"{tag}_{i}_loss_name": cfg["loss_name"],
                        f"{tag}_{i}_patch_size": cfg["patch_size"]})
        if tag != "b":
            wh, pad = T3.generate_patchinfo(hw[0], hw[1], psz, pst)
            out.update({f"{tag}_patch_wh_start": wh, f"{tag}_pad_info": np.array(pad)})
    # what DataLoader(dataset, 1) hands to run_iter for item 0 of (a) (shapes of the batched item)
    from torch.utils.data import DataLoader
    ds = T3.MVVidPatchDataset((16, 24), videos, (8, 12), (6, 8), poses, intr, loss_configs=cfgs)
    b = next(iter(DataLoader(ds, 1, shuffle=False)))
    out.update(batch_shapes=np.array([len(b[0].shape), len(b[1].shape), *b[2].shape, *b[3].shape, *b[4].shape]),
               batch_cfg_kinds=np.array([f"{k}:{'tensor' if torch.is_tensor(v) else 'list'}" for k, v in b[5].items()]))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, {t: int(out[t + "_len"]) for t in "abc"}, out["batch_cfg_kinds"])


def golden_stage1(name, H=24, W=40, D=4, hv=5, wv=7, seed=8, loop_mask=True, **arg_overrides):
    """SURVEY §8(f) N4, second half: the unmodified reference's stage-1 model (`MPI.MPMesh`, configs/mpi_base.txt) rendered
    and differentiated for one training view: rgbl, every extra term of MPI.py:596-652 and the gradients of
    `mean(rgbl * g_up) + sum_k extra_k * weight_k` w.r.t. `atlas` / `atlas_mask`."""
    import MPI  # reference
    args = ref_env.make_args(config="configs/mpi_base.txt", mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2,
                             mpi_h_scale=1.2, mpi_w_scale=1.2, learn_loop_mask=loop_mask, **arg_overrides)
    f = 0.8 * W
    ref_intrin = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(seed)
    m = MPI.MPMesh(args, H, W, np.eye(4, dtype=np.float32), ref_intrin, 1.0, 10.0)
    st, atlas_mask = MO.stage1_state(H, W, D, hv, wv, 2, 1.0, 10.0, 1.2, 1.2, seed=seed)
    assert torch.allclose(st.verts, m._verts.data) and torch.equal(st.faces, m.faces)
    assert torch.allclose(st.uvs, m.uvs.data, atol=1e-7) and tuple(st.atlas.shape) == tuple(m.atlas.shape)
    m.atlas.data = st.atlas.clone()
    if loop_mask:
        m.atlas_mask.data = atlas_mask.clone()
    ext, intr = _view(seed, H, W)
    m.train()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgbl, extra = m(H, W, ext, intr)
        _, var = m.render(H, W, ext @ m.ref_extrin[None].inverse(), intr)
    g = torch.Generator().manual_seed(seed + 3)
    g_up = torch.rand(rgbl.shape, generator=g) - 0.4
    weights = {k: float(getattr(args, k + "_loss_weight")) for k in extra}
    loss = (rgbl * g_up).mean()
    for k, v in extra.items():
        loss = loss + v.mean() * weights[k]
    loss.backward()
    out = dict(H=H, W=W, near=1.0, far=10.0, tar_extrin=ext, tar_intrin=intr, rgbl=rgbl, g_up=g_up, loss=loss.detach(),
               grad_atlas=m.atlas.grad, disp_norm=var["disp_norm"], alpha=var["alpha"], mpi=var["mpi"],
               blend_weight=var["blend_weight"], bg_color=str(args.bg_color), edge_scale=float(args.edge_scale),
               normalize_blendweight_fordepth=bool(args.normalize_blendweight_fordepth), loop_mask=bool(loop_mask),
               atlas_mask=atlas_mask, **{"extra_" + k: v.detach() for k, v in extra.items()},
               **{"w_" + k: w for k, w in weights.items()}, **_state_arrays(st))
    if loop_mask:
        out.update(grad_atlas_mask=m.atlas_mask.grad, loopmask3d=var["loopmask3d"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, {k: float(v) for k, v in extra.items()}, "loss", float(loss))


def golden_stage1_step(name, H=24, W=40, D=4, hv=5, wv=7, seed=14, lr=0.05):
    """One optimisation step of the stage-1 trainer: the body of `run_iter` (train_3d.py:189-236: scale-invariant MSE +
    loop-mask cross-entropy + weighted extra terms, backward, Adam) executed on the unmodified reference model."""
    import MPI  # reference
    args = ref_env.make_args(config="configs/mpi_base.txt", mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2,
                             mpi_h_scale=1.2, mpi_w_scale=1.2, add_intrin_noise=False, lrate=lr)
    f = 0.8 * W
    ref_intrin = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(seed)
    m = MPI.MPMesh(args, H, W, np.eye(4, dtype=np.float32), ref_intrin, 1.0, 10.0)
    st, atlas_mask = MO.stage1_state(H, W, D, hv, wv, 2, 1.0, 10.0, 1.2, 1.2, seed=seed)
    m.atlas.data, m.atlas_mask.data = st.atlas.clone(), atlas_mask.clone()
    ext, intr = _view(seed, H, W)
    pose = torch.inverse(ext)[:, :3, :]                                  # run_iter receives poses (train_3d.py:191)
    g = torch.Generator().manual_seed(seed + 1)
    b_rgbs = torch.rand(1, 3, H, W, generator=g)
    b_loopmask = (torch.rand(1, H, W, generator=g) > 0.5).float()
    opt = m.get_optimizer()
    import utils  # reference
    m.train()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgb, extra = m(H, W, utils.pose2extrin_torch(pose), intr)
    loop_mask = torch.clamp(rgb[:, -1], 0.001, 1 - 0.001)                # train_3d.py:201-212
    loop_loss = -(b_loopmask * torch.log(loop_mask) + (1 - b_loopmask) * torch.log(1 - loop_mask)).mean()
    rgb = rgb[:, :3]
    scale = torch.exp(torch.log((b_rgbs + 0.01) / (rgb.detach() + 0.01)).mean())   # train_3d.py:217-220
    rgb = rgb * ((scale + 3) / 4)
    img_loss = utils.img2mse(rgb, b_rgbs)
    loss = img_loss + loop_loss
    for k, v in extra.items():
        w = getattr(args, f"{k}_loss_weight")
        if w > 0:
            loss = loss + v.mean() * w
    opt.zero_grad()
    loss.backward()
    g_atlas, g_mask = m.atlas.grad.clone(), m.atlas_mask.grad.clone()
    opt.step()
    out = dict(H=H, W=W, lr=lr, pose=pose, tar_intrin=intr, rgb=b_rgbs, loopmask=b_loopmask, loss=loss.detach(),
               img_loss=img_loss.detach(), loop_loss=loop_loss.detach(), grad_atlas=g_atlas, grad_atlas_mask=g_mask,
               new_atlas=m.atlas.data, new_atlas_mask=m.atlas_mask.data, atlas_mask=atlas_mask, **_state_arrays(st))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, "loss", float(loss), "img", float(img_loss), "loop", float(loop_loss))


def golden_sparsify(name, H=32, W=48, D=4, hv=9, wv=13, seed=12):
    """Tile culling (`MPI.MPMesh.sparsify_faces`, MPI.py:289-442) of the unmodified reference on a stage-1 model whose
    alpha / loop-mask logits are blobs on the untouched initial value: the state dict after culling (what stage 2
    loads) and a render of the culled model."""
    import MPI  # reference
    args = ref_env.make_args(config="configs/mpi_base.txt", mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2,
                             mpi_h_scale=1.0, mpi_w_scale=1.0)
    f = 0.8 * W
    ref_intrin = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(seed)
    m = MPI.MPMesh(args, H, W, np.eye(4, dtype=np.float32), ref_intrin, 1.0, 10.0)
    g = torch.Generator().manual_seed(seed)
    Ha, Wa = m.atlas.shape[-2:]
    atlas = m.atlas.data.clone()
    mask = m.atlas_mask.data.clone()
    for _ in range(40):                                             # blobs of real content on the -3 background
        y, x = int(torch.randint(0, Ha - 6, (1,), generator=g)), int(torch.randint(0, Wa - 8, (1,), generator=g))
        hh, ww = int(torch.randint(3, 7, (1,), generator=g)), int(torch.randint(4, 9, (1,), generator=g))
        atlas[:, 3, y:y + hh, x:x + ww] = torch.randn((hh, ww), generator=g) + 1.0
        if torch.rand(1, generator=g) < 0.5:
            mask[:, 0, y:y + hh, x:x + ww] = torch.randn((hh, ww), generator=g) + 2.0
    m.atlas.data, m.atlas_mask.data = atlas.clone(), mask.clone()
    m.sparsify_faces(erode_num=1, alpha_thresh=0.05)
    sd = m.state_dict()
    ext, intr = _view(seed, H, W)
    m.eval()
    with torch.no_grad():
        rgbl, _ = m(H, W, ext, intr)
    key = lambda k: k.replace("self.", "self_")
    out = {"sd_" + key(k): v for k, v in sd.items()}
    out.update(H=H, W=W, D=D, hv=hv, wv=wv, atlas0=atlas, atlas_mask0=mask, erode_num=1, alpha_thresh=0.05, tar_extrin=ext,
               tar_intrin=intr, rgbl=rgbl)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **_np(out))
    print(name, {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in sd.items() if "atlas" in k or "faces" in k})


def main():
    assert ref_env.reference_available(), "needs /root/reference"
    ref_env.enable()
    os.makedirs(OUT, exist_ok=True)
    golden_pointwise("pointwise")
    golden_nnerr("nnerr")
    golden_lod("lod")
    golden_ckpt("ckpt")
    golden_dataset("dataset")
    golden_render("render_dense", "dense", seed=0)
    golden_render("render_sparse", "sparse", seed=1, D=6, hv=6, wv=9, T=2)
    golden_render_bg_random("render_bg_random")
    golden_step("step_dense_refcfg", "dense", LOSS_CFG_REF, seed=2)
    golden_step("step_sparse_othercfg", "sparse", LOSS_CFG_OTHER, seed=3, D=6, hv=6, wv=9)
    # the optional terms of MPV.py:455-466, 511-515, 533-551 (weight 0 / off in the shipped stage-2 configs)
    terms = dict(sparsity_loss_weight=0.004, density_loss_weight=0.02, d_smooth_loss_weight=0.1)
    golden_step("step_dense_terms", "dense", LOSS_CFG_REF, seed=4, args=dict(bg_color="0.2#0.5#0.9", **terms))
    golden_step("step_sparse_terms", "sparse", LOSS_CFG_OTHER, seed=6, D=6, hv=6, wv=9, args=terms)   # (seed 5 has an NN near-tie)
    golden_sparsify("stage1_sparsify")
    golden_stage1_step("stage1_step")
    golden_stage1("stage1_loopmask", d_smooth_loss_weight=0.1, l_smooth_loss_weight=0.05, edge_scale=0.5)
    golden_stage1("stage1_bg_normdepth", seed=9, loop_mask=False, d_smooth_loss_weight=0.1, bg_color="0.9#0.1#0.4",
                  normalize_blendweight_fordepth=True, edge_scale=0.5)
    golden_loss("loss_lm_alpha0", "Patch3DGPNNLowMemLoss", macro_block=15, patch_size=5, stride=2, patcht_size=3,
                stridet=1, rou="-2", scaling=0.1, alpha=0.0)
    golden_loss("loss_lm_noalpha", "Patch3DGPNNLowMemLoss", macro_block=11, patch_size=3, stride=2, patcht_size=3,
                stridet=1, rou="-2", scaling=0.1, alpha=10000.0, smooth=True)
    golden_loss("loss_direct_p7", "Patch3DGPNNDirectLoss", patch_size=7, stride=4, patcht_size=2, stridet=2,
                rou="0", scaling=0.2, alpha=0.5, t=9, h=23, w=27)
    golden_loss("loss_lm_abs", "Patch3DGPNNLowMemLoss", macro_block=64, patch_size=3, stride=1, patcht_size=1,
                stridet=1, rou="abs", scaling=0.2, alpha=10000.0, t=4, F=5, h=9, w=10)


if __name__ == "__main__":
    main()
