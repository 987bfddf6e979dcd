"""CPU: host-side logic (quad table, homographies, loss descriptors, args) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import mpv_oracle as MO
from util import load_golden, state_from_golden
from videoloop3d_b200 import _lib, build, ops, tiles

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emulate_kernel_geometry(st, table, homs, cx, cy, H, W):
    """float32 re-statement of composite.cu's plane_grid()/make_taps() address maths (test only)."""
    D, qh, qw = st.mpi_d, st.hv - 1, st.wv - 1
    r, c = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u = (c + 0.5 - cx).astype(np.float32).reshape(-1)
    v = (r + 0.5 - cy).astype(np.float32).reshape(-1)
    out = []
    for d in range(D):
        h = homs[d]
        w_ = h[6] * u + h[7] * v + h[8]
        with np.errstate(divide="ignore", invalid="ignore"):
            gx = (h[0] * u + h[1] * v + h[2]) / w_
            gy = (h[3] * u + h[4] * v + h[5]) / w_
        ok = (w_ > 0) & (gx > 0) & (gx < qw) & (gy > 0) & (gy < qh)
        qx = np.clip(np.nan_to_num(gx).astype(np.int32), 0, qw - 1)
        qy = np.clip(np.nan_to_num(gy).astype(np.int32), 0, qh - 1)
        e = table[(d * qh + qy) * qw + qx]
        ok &= e["kind"] != 0
        ax = e["x0i"] + (e["x0f"] + (gx - qx) * e["sx"])
        ay = e["y0i"] + (e["y0f"] + (gy - qy) * e["sy"])
        out.append((ok, e["kind"], ax, ay))
    return out


@pytest.mark.parametrize("name", ["render_dense", "render_sparse"])
def test_quad_table_and_homographies_match_oracle_geometry(name):
    g = load_golden(name)
    st = state_from_golden(g)
    H, W, D, hv, wv = int(g["H"]), int(g["W"]), st.mpi_d, st.hv, st.wv
    grids = tiles.plane_grids(st.verts.numpy(), D, hv, wv)
    table = tiles.build_quad_table(D, hv, wv, st.faces.numpy(), st.uvs.numpy(), st.uvfaces.numpy(),
                                   tuple(st.atlas.shape[-2:]), st.faces_dyn.numpy(), st.uvs_dyn.numpy(),
                                   st.uvfaces_dyn.numpy(), tuple(st.atlas_dyn.shape[-2:]))
    assert table.dtype.itemsize == ctypes.sizeof(_lib.Quad) == 32
    ext = g["tar_extrin"].reshape(4, 4).astype(np.float64) @ np.linalg.inv(st.ref_extrin.numpy().astype(np.float64))
    homs, cx, cy = tiles.view_homographies(grids, hv - 1, wv - 1, ext, g["tar_intrin"], np.eye(4), H, W)
    geo = MO.geometry(st, H, W, torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"]))
    for d, (ok, kind, ax, ay) in enumerate(_emulate_kernel_geometry(st, table, homs, cx, cy, H, W)):
        oh = geo["hit"][:, d].numpy()
        assert np.array_equal(ok, oh)
        assert np.array_equal(kind[oh], geo["kind"][:, d].numpy()[oh])
        assert np.abs(ax[oh] - geo["ax"][:, d].numpy()[oh]).max(initial=0) < 1e-4
        assert np.abs(ay[oh] - geo["ay"][:, d].numpy()[oh]).max(initial=0) < 1e-4


def test_geometry_errors():
    st = MO.dense_state(16, 24, 2, 3, 4, 1, 1, 1.0, 10.0)
    v = st.verts.numpy().copy()
    v[5, 0] += 0.3
    with pytest.raises(tiles.GeometryError):
        tiles.plane_grids(v, 2, 3, 4)
    grids = tiles.plane_grids(st.verts.numpy(), 2, 3, 4)
    behind = np.eye(4)
    behind[2, 3] = -5.0          # camera origin at z = +5: inside the plane stack
    with pytest.raises(tiles.GeometryError):
        tiles.view_homographies(grids, 2, 3, behind, np.array([[20., 0, 12], [0, 20, 8], [0, 0, 1]]), np.eye(4), 16, 24)
    bad_faces = st.faces_dyn.numpy().copy()
    bad_faces[0, 1] += 1
    with pytest.raises(tiles.GeometryError):
        tiles.build_quad_table(2, 3, 4, st.faces.numpy(), st.uvs.numpy(), st.uvfaces.numpy(), (4, 4), bad_faces,
                               st.uvs_dyn.numpy(), st.uvfaces_dyn.numpy(), tuple(st.atlas_dyn.shape[-2:]))


def test_loss_desc_fitting_matches_reference_rules():
    d = ops.make_loss_desc((50, 3, 180, 320), (1, 1, 1), (258, 3, 180, 320), (1, 1, 1), 11, 3, 4, 1, 0.0)
    assert (d.t, d.h, d.w) == (50, 179, 319) and (d.n1, d.n2, d.ho, d.wo) == (48, 256, 43, 78)
    assert d.use_alpha == 1 and d.alpha == 0.0
    d = ops.make_loss_desc((50, 3, 180, 320), (1, 1, 1), (64, 3, 180, 320), (1, 1, 1), 3, 3, 2, 1, 10000.0)
    assert (d.h, d.w, d.ho, d.wo) == (179, 319, 89, 159) and d.use_alpha == 0
    d = ops.make_loss_desc((9, 3, 23, 27), (1, 1, 1), (12, 3, 23, 27), (1, 1, 1), 7, 2, 4, 2, 0.5, fit=False)
    assert (d.t, d.h, d.w, d.n1, d.n2, d.ho, d.wo) == (9, 23, 27, 4, 6, 5, 6)
    with pytest.raises(ValueError):
        ops.make_loss_desc((2, 3, 23, 27), (1, 1, 1), (12, 3, 23, 27), (1, 1, 1), 7, 3, 4, 1, 0.5)
    assert ops.parse_rou("-2") == (0, -2.0) and ops.parse_rou("mse")[0] == 1 and ops.parse_rou("abs")[0] == 2


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads (no GPU needed) and exports exactly what include/vl3d.h declares."""
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "vl3d.h")).read()
    declared = set(re.findall(r"\b(vl3d_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vl3d_version() == 100
    assert ctypes.sizeof(_lib.View) == 4 * 9 + 4 * 2 + 4 * 32 * 9 + 4       # ... + flags
    assert ctypes.sizeof(_lib.LossDesc) == 4 * 12 + 8 * 6 + 8


def test_argument_validation_happens_before_any_launch():
    lib = _lib.load()
    v = _lib.View()
    v.D = 64
    with pytest.raises(_lib.Vl3dError, match="D=64"):
        _lib.call("vl3d_composite_fwd", ctypes.byref(v), ctypes.c_void_p(16), None, None, None, 1, 0,
                  ctypes.c_void_p(16), None, None, None, None, None)
    d = ops.make_loss_desc((50, 3, 180, 320), (1, 1, 1), (258, 3, 180, 320), (1, 1, 1), 11, 3, 4, 1, 0.0)
    with pytest.raises(_lib.Vl3dError, match="NULL"):
        _lib.call("vl3d_patchnn_search", ctypes.byref(d), None, None, 0, 1, None, None)
    with pytest.raises(_lib.Vl3dError, match="step"):
        _lib.call("vl3d_adam_step", ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16),
                  8, 0, 0.1, 0.9, 0.999, 6e-8, None)
    assert b"step" in lib.vl3d_last_error_string()
    # the optional-terms entry points (csrc/terms.cu): every check precedes the launch, so no device is needed
    P16 = ctypes.c_void_p(16)
    v = _lib.View()
    v.H, v.W, v.D, v.qh, v.qw = 8, 8, 2, 1, 1
    coef = (ctypes.c_float * 6)()
    with pytest.raises(_lib.Vl3dError, match="no output"):
        _lib.call("vl3d_composite_terms_fwd", ctypes.byref(v), P16, None, None, None, 1, coef, 1e-4, None, None, None, None)
    with pytest.raises(_lib.Vl3dError, match="inv_depth_host"):
        _lib.call("vl3d_composite_terms_fwd", ctypes.byref(v), P16, None, None, None, 1, None, 1e-4, None, P16, None, None)
    with pytest.raises(_lib.Vl3dError, match="sparsity_eps"):
        _lib.call("vl3d_composite_terms_fwd", ctypes.byref(v), P16, None, None, None, 1, coef, 0.0, P16, None, None, None)
    with pytest.raises(_lib.Vl3dError, match="bad T"):
        _lib.call("vl3d_composite_terms_fwd", ctypes.byref(v), P16, None, None, None, 0, coef, 1e-4, P16, None, None, None)
    with pytest.raises(_lib.Vl3dError, match="inv_depth_host"):
        _lib.call("vl3d_composite_terms_bwd", ctypes.byref(v), P16, None, None, None, 1, None, 1e-4, None, P16, None, None, None, None)
    with pytest.raises(_lib.Vl3dError, match="16-byte"):
        _lib.call("vl3d_composite_terms_bwd", ctypes.byref(v), P16, None, None, None, 1, coef, 1e-4, P16, None, None,
                  ctypes.c_void_p(8), None, None)
    # nothing upstream: a no-op that returns success without a launch
    assert _lib.call("vl3d_composite_terms_bwd", ctypes.byref(v), P16, None, None, None, 1, coef, 1e-4, None, None, None,
                     None, None, None) == 0


def test_product_refuses_cpu_tensors():
    from videoloop3d_b200 import Patch3DGPNNLowMemLoss, Vl3dError
    x, y = torch.rand(1, 3, 5, 9, 9), torch.rand(1, 3, 6, 9, 9)
    with pytest.raises(Vl3dError):
        Patch3DGPNNLowMemLoss()(x, y, patch_size=3, stride=2, patcht_size=3, stridet=1)


def _lod_model(g, dense=False):
    from videoloop3d_b200.testing import model_from_tensors
    from videoloop3d_b200 import MPMeshVid, default_args
    H, W, tile = int(g["H"]), int(g["W"]), int(g["tile"])
    if dense:
        args = default_args(mpi_d=int(g["mpi_d"]), mpi_h_verts=int(g["hv"]), mpi_w_verts=int(g["wv"]),
                            atlas_grid_h=int(g["dense_grid_h"]), mpv_frm_num=g["dense_atlas_dyn"].shape[0],
                            mpi_h_scale=float(g["dense_scale"]), mpi_w_scale=float(g["dense_scale"]))
        f = 0.8 * W
        m = MPMeshVid(args, H, W, np.eye(4, dtype=np.float32),
                      np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32), 1.0, 10.0)
        m.atlas_dyn.data = torch.as_tensor(g["dense_atlas_dyn"]).clone()
        return m
    m = model_from_tensors(g, H, W, torch.device("cpu"))
    m.atlas_grid_h, m.atlas_grid_w = g["atlas"].shape[-2] // tile, g["atlas"].shape[-1] // tile
    m.atlas_full_h, m.atlas_full_w = g["atlas"].shape[-2:]
    m.atlas_grid_dyn_h, m.atlas_grid_dyn_w = g["atlas_dyn"].shape[-2] // tile, g["atlas_dyn"].shape[-1] // tile
    m.atlas_full_dyn_h, m.atlas_full_dyn_w = g["atlas_dyn"].shape[-2:]
    return m


def test_lod_matches_reference():
    """`MPMeshVid.lod` against the unmodified reference's (MPV.py:140-198;
    oracle/make_golden.py::golden_lod): per-tile bilinear resize + uv re-alignment of a tile-culled model,
    full -> 0.5 -> 1.0, and the dense layout.  The down-sampling golden uses torchvision's behaviour at the
    reference's pinned version (no anti-aliasing for tensors); up-sampling does not depend on it."""
    from util import load_golden
    g = load_golden("lod")
    m = _lod_model(g)
    m.lod(0.5)
    for k in ("atlas", "atlas_dyn", "uvs", "uvs_dyn"):
        got, ref = getattr(m, k).data, torch.as_tensor(g["half_" + k])
        assert tuple(got.shape) == tuple(ref.shape), k
        assert float((got - ref).abs().max()) < 2e-6, k
    assert m.atlas_dyn.data.is_contiguous(memory_format=torch.channels_last)
    m.lod(1.0)
    for k in ("atlas", "atlas_dyn", "uvs", "uvs_dyn"):
        got, ref = getattr(m, k).data, torch.as_tensor(g["full_" + k])
        assert tuple(got.shape) == tuple(ref.shape), k
        assert float((got - ref).abs().max()) < 2e-6, k
    # up-sampling alone, from the reference's anti-aliased half-size state: independent of the torchvision default
    m2 = _lod_model(g)
    m2.lod(0.5)
    for k in ("atlas", "atlas_dyn"):
        getattr(m2, k).data.copy_(torch.as_tensor(g["half_aa_" + k]))
    m2.lod(1.0)
    for k in ("atlas", "atlas_dyn", "uvs", "uvs_dyn"):
        assert float((getattr(m2, k).data - torch.as_tensor(g["full_aa_" + k])).abs().max()) < 2e-6, k
    md = _lod_model(g, dense=True)
    md.lod(0.5)
    assert float((md.atlas_dyn.data - torch.as_tensor(g["dense_half_atlas_dyn"])).abs().max()) < 2e-6
    md.lod(1.0)
    assert float((md.atlas_dyn.data - torch.as_tensor(g["dense_full_atlas_dyn"])).abs().max()) < 2e-6


def test_checkpoint_format_matches_reference():
    """a stage-1 checkpoint goes through `init_from_mpi` -> `lod(0.5)` -> `lod(1.0)` ->
    `state_dict()` exactly as in the unmodified reference (oracle/make_golden.py::golden_ckpt): same keys,
    same scalars, same tensors (the atlases are stored in the reference's logical NCHW shape whatever the
    memory format), and the static-only branch of `init_from_mpi`."""
    from util import ckpt_dict, ckpt_model, load_golden
    g = load_golden("ckpt")
    m = ckpt_model(g, torch.device("cpu"))
    m.init_from_mpi(ckpt_dict(g, "stage1_"))
    assert m.atlas_dyn.shape[0] == int(g["T"]) and m.frm_num == int(g["T"])
    m.atlas_dyn.data = m.atlas_dyn.data + torch.as_tensor(g["noise"])
    m.lod(0.5)
    m.lod(1.0)
    sd, ref = m.state_dict(), ckpt_dict(g, "sd_")
    assert set(sd.keys()) == set(ref.keys())
    for k, v in ref.items():
        if torch.is_tensor(v):
            assert tuple(sd[k].shape) == tuple(v.shape), k
            assert float((sd[k].double() - v.double()).abs().max()) < 2e-6, k
        else:
            assert sd[k] == v, k
    m3 = ckpt_model(g, torch.device("cpu"))
    m3.init_from_mpi({k: v for k, v in ckpt_dict(g, "stage1_").items() if "dyn" not in k})
    sd3, ref3 = m3.state_dict(), ckpt_dict(g, "sd3_")
    assert set(sd3.keys()) == set(ref3.keys())
    for k, v in ref3.items():
        if torch.is_tensor(v):
            assert tuple(sd3[k].shape) == tuple(v.shape), k
            assert torch.allclose(sd3[k].double(), v.double(), rtol=0, atol=1e-7), k
        else:
            assert sd3[k] == v, k


def test_dataset_matches_reference():
    """`MVVidPatchDataset` / `generate_patchinfo` give the unmodified reference's items
    (train_3dvid.py:22-66, utils.py:115-134; oracle/make_golden.py::golden_dataset): patch origins and order,
    shifted intrinsics, crops of the resized + padded videos bit for bit, loss config per view — and `batches()`
    yields what `DataLoader(dataset, 1)` yields."""
    from util import load_golden
    from videoloop3d_b200 import MVVidPatchDataset, generate_patchinfo
    g = load_golden("dataset")
    V = int(g["V"])
    videos = [g[f"video{i}"] for i in range(V)]
    poses, intr = torch.as_tensor(g["poses"]), torch.as_tensor(g["intrins"])
    cfgs = [dict(loss_name="gpnn_lm", patch_size=5), dict(loss_name="gpnn_lm", patch_size=3)]
    for tag in "abc":
        hw, psz, pst = tuple(g[f"{tag}_hw"]), tuple(g[f"{tag}_patch_size"]), tuple(g[f"{tag}_patch_stride"])
        ds = MVVidPatchDataset(hw, videos, psz, pst, poses, intr, loss_configs=cfgs)
        assert len(ds) == int(g[f"{tag}_len"])
        for i in range(len(ds)):
            w0, h0, pose, k, crops, cfg = ds[i]
            assert [int(w0), int(h0)] == g[f"{tag}_{i}_wh"].tolist()
            assert torch.equal(pose, torch.as_tensor(g[f"{tag}_{i}_pose"]))
            assert torch.equal(k, torch.as_tensor(g[f"{tag}_{i}_intrin"])) and k.dtype == torch.float32
            assert torch.equal(crops, torch.as_tensor(g[f"{tag}_{i}_crops"]))
            assert cfg["patch_size"] == int(g[f"{tag}_{i}_patch_size"])
            cfg["patch_size"] = -1                                        # items own a deep copy of the config
        assert cfgs[0]["patch_size"] == 5
        if tag != "b":
            wh, pad = generate_patchinfo(hw[0], hw[1], psz, pst)
            assert torch.equal(wh, torch.as_tensor(g[f"{tag}_patch_wh_start"])) and list(pad) == g[f"{tag}_pad_info"].tolist()
    ds = MVVidPatchDataset((16, 24), videos, (8, 12), (6, 8), poses, intr, loss_configs=cfgs, pin_memory=False)
    b = next(iter(ds.batches(shuffle=False)))
    shapes = [b[0].dim(), b[1].dim(), *b[2].shape, *b[3].shape, *b[4].shape]
    assert shapes == g["batch_shapes"].tolist()
    assert b[5]["loss_name"] == ["gpnn_lm"] and torch.is_tensor(b[5]["patch_size"]) and int(b[5]["patch_size"][0]) == 5
    from torch.utils.data import DataLoader
    ds.loss_configs = [dict(loss_name="gpnn_lm", patch_size=5, scaling=0.1, alpha=0.0), dict(loss_name="gpnn_lm", patch_size=3)]
    ours, theirs = next(iter(ds.batches(shuffle=False)))[5], next(iter(DataLoader(ds, 1, shuffle=False)))[5]
    assert set(ours) == set(theirs)
    for k in ours:                                                        # same types, dtypes and values as default_collate
        if torch.is_tensor(theirs[k]):
            assert ours[k].dtype == theirs[k].dtype and torch.equal(ours[k], theirs[k]), k
        else:
            assert ours[k] == theirs[k], k
    assert ours["scaling"][0].item() == 0.1
    assert b[4].untyped_storage().data_ptr() == ds.videos[0].untyped_storage().data_ptr()     # a view, not a copy
    seen = sorted((int(x[0]), int(x[1]), float(x[3][0, 0, 2])) for x in ds.batches(shuffle=True, generator=torch.Generator().manual_seed(1)))
    assert len(seen) == len(ds) and len(set(seen)) > 1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host CPU) prints ONE JSON line with the keys the driver
    reads; the synthetic target is the same 8-bit video as bytes and as floats."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.WORKLOADS["tiny"]
    f, b = bench.make_target(wl, None), bench.make_target(wl, None, as_bytes=True)
    assert b.dtype == torch.uint8 and torch.equal(f, b.float() / 255)


@pytest.mark.parametrize("seed", range(6))
def test_host_geometry_matches_oracle_on_random_views_and_layouts(seed):
    """The quad table + per-plane homographies the host hands to the kernels (tiles.py) reproduce the oracle's
    per-pixel geometry (hit mask exactly, tile kind, atlas coordinates to 1e-4 texel) for seeded random tile
    cullings, atlas scales, rotations about all three axes, translations and principal-point shifts — including the
    patch views of the dataset (principal point moved by a patch origin, utils.py:196-200)."""
    rng = np.random.default_rng(100 + seed)
    H, W = int(rng.integers(20, 40)), int(rng.integers(28, 56))
    D, hv, wv = int(rng.integers(2, 7)), int(rng.integers(3, 7)), int(rng.integers(3, 8))
    scale = float(rng.uniform(1.0, 1.4))
    if seed % 2:
        st = MO.sparse_state(H, W, D, hv, wv, 2, 1.0, 10.0, tile=int(rng.integers(3, 9)), occupancy=float(rng.uniform(0.3, 0.9)),
                             dyn_frac=float(rng.uniform(0.2, 0.8)), h_scale=scale, w_scale=scale, seed=seed)
    else:
        st = MO.dense_state(H, W, D, hv, wv, 1, 2, 1.0, 10.0, scale, scale, seed=seed)
        st.atlas = st.atlas[:, :, :1, :1].clone()
    ang = rng.uniform(-0.06, 0.06, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    ext = np.eye(4, dtype=np.float32)
    ext[:3, :3] = Rx @ Ry @ Rz
    ext[:3, 3] = rng.uniform(-0.1, 0.1, 3) * [1, 1, 0.3]
    f = 0.8 * W * float(rng.uniform(0.8, 1.3))
    intr = np.array([[f, 0, W / 2 + rng.uniform(-6, 6)], [0, f, H / 2 + rng.uniform(-6, 6)], [0, 0, 1]], dtype=np.float32)
    grids = tiles.plane_grids(st.verts.numpy(), D, hv, wv)
    table = tiles.build_quad_table(D, hv, wv, st.faces.numpy(), st.uvs.numpy(), st.uvfaces.numpy(),
                                   tuple(st.atlas.shape[-2:]), st.faces_dyn.numpy(), st.uvs_dyn.numpy(),
                                   st.uvfaces_dyn.numpy(), tuple(st.atlas_dyn.shape[-2:]))
    rel = ext.astype(np.float64) @ np.linalg.inv(st.ref_extrin.numpy().astype(np.float64))
    homs, cx, cy = tiles.view_homographies(grids, hv - 1, wv - 1, rel, intr[None], np.eye(4), H, W)
    geo = MO.geometry(st, H, W, torch.as_tensor(ext)[None], torch.as_tensor(intr)[None])
    n_hit = 0
    for d, (ok, kind, ax, ay) in enumerate(_emulate_kernel_geometry(st, table, homs, cx, cy, H, W)):
        oh = geo["hit"][:, d].numpy()
        # a pixel whose ray passes within fp32 rounding of a quad edge may legitimately fall on either side
        gx_edge = np.zeros_like(oh)
        if not np.array_equal(ok, oh):
            diff = ok != oh
            assert diff.sum() <= 2, (d, int(diff.sum()))
            gx_edge = diff
        keep = oh & ~gx_edge & ok
        n_hit += int(keep.sum())
        assert np.array_equal(kind[keep], geo["kind"][:, d].numpy()[keep])
        assert np.abs(ax[keep] - geo["ax"][:, d].numpy()[keep]).max(initial=0) < 1e-4
        assert np.abs(ay[keep] - geo["ay"][:, d].numpy()[keep]).max(initial=0) < 1e-4
    assert n_hit > H * W // 4
    # the per-plane inverse view depth the terms kernels use (tiles.view_inv_depth): exactly linear in the pixel, equal to
    # 1 / (the oracle's view depth = pytorch3d's zbuf); an affine normalisation folds into the coefficients
    coef = tiles.view_inv_depth(grids, rel, intr[None], np.eye(4), H, W).astype(np.float64)
    coef_n = tiles.view_inv_depth(grids, rel, intr[None], np.eye(4), H, W, scale=1.0 / 0.9, offset=-0.1 / 0.9).astype(np.float64)
    rr, cc = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u, v = (cc + 0.5 - W / 2.0).reshape(-1), (rr + 0.5 - H / 2.0).reshape(-1)
    for d in range(D):
        oh = geo["hit"][:, d].numpy()
        inv_ref = 1.0 / geo["depth"][:, d].numpy()[oh]
        got = coef[d, 0] * u[oh] + coef[d, 1] * v[oh] + coef[d, 2]
        assert np.abs(got - inv_ref).max(initial=0) < 2e-6 * np.abs(inv_ref).max(initial=1)
        got_n = coef_n[d, 0] * u[oh] + coef_n[d, 1] * v[oh] + coef_n[d, 2]
        assert np.abs(got_n - (inv_ref - 0.1) / 0.9).max(initial=0) < 2e-6


def test_argument_validation_of_the_data_entry_points():
    """vl3d_u8_to_unit / vl3d_to8b / vl3d_patch_l1 / vl3d_vote_loss reject bad arguments with a VL3D_E* code and a
    message before touching the device (runs without a GPU)."""
    lib = _lib.load()
    p16 = ctypes.c_void_p(16)
    with pytest.raises(_lib.Vl3dError, match="NULL"):
        _lib.call("vl3d_u8_to_unit", None, p16, 3, 4, 4, 16, 4, None)
    with pytest.raises(_lib.Vl3dError, match="strides"):
        _lib.call("vl3d_u8_to_unit", p16, p16, 3, 4, 8, 16, 4, None)          # row stride < width
    with pytest.raises(_lib.Vl3dError, match="strides"):
        _lib.call("vl3d_u8_to_unit", p16, p16, 3, 4, 4, 8, 4, None)           # plane stride < one image
    with pytest.raises(_lib.Vl3dError, match="to8b"):
        _lib.call("vl3d_to8b", p16, p16, 0, 4, 4, None)
    d = ops.make_loss_desc((50, 3, 180, 320), (1, 1, 1), (258, 3, 180, 320), (1, 1, 1), 11, 3, 4, 1, 0.0)
    with pytest.raises(_lib.Vl3dError, match="NULL"):
        _lib.call("vl3d_vote_loss", ctypes.byref(d), None, None, None, None, 0, -2.0, 0.1, 1.0, 50, 180, 320, 0, 50,
                  0, 180, 0, None, None, None, None, None, None)
    with pytest.raises(_lib.Vl3dError, match="frame range"):
        _lib.call("vl3d_vote_loss", ctypes.byref(d), p16, None, p16, p16, 0, -2.0, 0.1, 1.0, 50, 180, 320, 7, 3,
                  0, 180, 0, None, None, None, p16, p16, None)
    with pytest.raises(_lib.Vl3dError, match="row range"):
        _lib.call("vl3d_vote_loss", ctypes.byref(d), p16, None, p16, p16, 0, -2.0, 0.1, 1.0, 50, 180, 320, 0, 50,
                  90, 181, 0, None, None, None, p16, p16, None)
    with pytest.raises(_lib.Vl3dError, match="row range"):
        _lib.call("vl3d_scale_log_sum", p16, 4, p16, 4, 8, 8, 5, 3, p16, p16, None)
    assert lib.vl3d_vote_partials(50, 180, 320) == 50 * 10 * 23 and lib.vl3d_vote_partials(0, 1, 1) == 0


def test_constructor_state_equals_reference_state():
    """The dense model `MPMeshVid.__init__` builds (from the quad grid: tiles.quad_grid_faces / dense_atlas_uvs) carries
    exactly the tensors the unmodified reference's constructor produced for the same arguments (MPV.py:26-104; stored in
    tests/golden/render_dense.npz by oracle/make_golden.py::golden_render) — checkpoints are interchangeable."""
    from util import load_golden
    from videoloop3d_b200 import MPMeshVid, default_args
    g = load_golden("render_dense")
    H, W, D, hv, wv = int(g["H"]), int(g["W"]), int(g["mpi_d"]), int(g["hv"]), int(g["wv"])
    args = default_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpv_frm_num=int(g["T"]), mpi_h_scale=1.2,
                        mpi_w_scale=1.2)
    f = 0.8 * W
    m = MPMeshVid(args, H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32),
                  1.0, 10.0)
    assert torch.equal(m.faces_dyn, torch.as_tensor(g["faces_dyn"]).long()) and torch.equal(m.uvfaces_dyn, m.faces_dyn)
    assert len(m.faces) == 0 and len(m.uvfaces) == 0 and len(m.uvs) == 0
    assert torch.equal(m.uvs_dyn.data, torch.as_tensor(g["uvs_dyn"]))            # bit for bit
    assert torch.allclose(m._verts.data, torch.as_tensor(g["verts"]), rtol=0, atol=0)
    assert torch.equal(m.planedepth, torch.as_tensor(g["planedepth"]))
    assert tuple(m.atlas_dyn.shape) == tuple(g["atlas_dyn"].shape)
    assert set(m.losses) == {"swd", "gpnn", "gpnn_lm", "mse", "avg"}
    sd = m.state_dict()
    for k in ("self.is_sparse", "self.atlas_full_w", "self.atlas_full_h", "self.atlas_grid_h", "self.atlas_grid_w", "self.has_dyn",
              "self.atlas_full_dyn_w", "self.atlas_full_dyn_h", "self.atlas_grid_dyn_h", "self.atlas_grid_dyn_w"):
        assert k in sd


def test_stage1_constructor_state_equals_reference_state():
    """`MPMesh.__init__` (second half) carries the tensors the unmodified reference's stage-1 constructor
    produced for the same arguments (MPI.py:38-124; stored in tests/golden/stage1_loopmask.npz by
    oracle/make_golden.py::golden_stage1, which asserts them against the reference's own)."""
    from util import load_golden
    from videoloop3d_b200 import MPMesh, default_args_stage1
    g = load_golden("stage1_loopmask")
    H, W, D, hv, wv = int(g["H"]), int(g["W"]), int(g["mpi_d"]), int(g["hv"]), int(g["wv"])
    args = default_args_stage1(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpi_h_scale=1.2, mpi_w_scale=1.2)
    f = 0.8 * W
    torch.manual_seed(0)
    m = MPMesh(args, H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32),
               1.0, 10.0)
    assert {k for k, _ in m.named_parameters()} == {"_verts", "uvs", "atlas", "atlas_mask"}
    assert {k for k, _ in m.named_buffers()} == {"ref_extrin", "ref_intrin", "planedepth", "uvfaces", "faces"}
    assert torch.equal(m.faces, torch.as_tensor(g["faces"]).long()) and torch.equal(m.uvfaces, m.faces)
    assert torch.equal(m.uvs.data, torch.as_tensor(g["uvs"]))                    # bit for bit
    assert torch.equal(m._verts.data, torch.as_tensor(g["verts"]))
    assert torch.equal(m.planedepth, torch.as_tensor(g["planedepth"]))
    assert tuple(m.atlas.shape) == tuple(g["atlas"].shape) and tuple(m.atlas_mask.shape) == tuple(g["atlas_mask"].shape)
    # the constructor's only RNG draw is the atlas (MPI.py:102), alpha / mask logits start at -3 (MPI.py:35,103,118)
    torch.manual_seed(0)
    assert torch.equal(m.atlas.data[:, :3], torch.rand((1, 4) + tuple(m.atlas.shape[-2:]))[:, :3])
    assert float(m.atlas.data[:, 3].min()) == float(m.atlas.data[:, 3].max()) == -3.0 and float(m.atlas_mask.data.max()) == -3.0
    assert not m.is_sparse and not m.has_dyn
    m.eval()
    with pytest.raises(Exception, match="CUDA"):                                 # no CPU fallback
        m.render(H, W, np.eye(4)[None], torch.as_tensor(g["tar_intrin"]))


def test_stage1_tile_culling_matches_reference():
    """`MPMesh.sparsify_faces` + `state_dict` (MPI.py:289-442, 207-221): the same quads kept / made dynamic, the same packed
    atlases, uv corners and layout scalars as the unmodified reference (tests/golden/stage1_sparsify.npz), i.e. the
    checkpoint stage 2 loads.  Pure host / torch logic: runs on the CPU here."""
    from util import ckpt_dict, load_golden
    from videoloop3d_b200 import MPMesh, default_args_stage1
    g = load_golden("stage1_sparsify")
    H, W, D, hv, wv = (int(g[k]) for k in ("H", "W", "D", "hv", "wv"))
    args = default_args_stage1(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpi_h_scale=1.0, mpi_w_scale=1.0)
    f = 0.8 * W
    m = MPMesh(args, H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32),
               1.0, 10.0)
    m.atlas.data = torch.as_tensor(g["atlas0"]).clone()
    m.atlas_mask.data = torch.as_tensor(g["atlas_mask0"]).clone()
    info = m.sparsify_faces(erode_num=int(g["erode_num"]), alpha_thresh=float(g["alpha_thresh"]))
    ref = ckpt_dict(g, "sd_")
    sd = m.state_dict()
    assert set(sd) == set(ref)
    assert info["kept"] == (len(ref["faces"]) + len(ref["faces_dyn"])) // 2 and info["dynamic"] == len(ref["faces_dyn"]) // 2
    for k in ("faces", "faces_dyn", "uvfaces", "uvfaces_dyn"):
        assert torch.equal(sd[k], ref[k].long()), k
    for k in ("uvs", "uvs_dyn", "_verts", "planedepth", "ref_extrin", "ref_intrin"):
        assert torch.allclose(sd[k], ref[k], rtol=0, atol=1e-6), k
    for k in ("atlas", "atlas_dyn"):
        assert tuple(sd[k].shape) == tuple(ref[k].shape) and float((sd[k] - ref[k]).abs().max()) < 1e-5, k
    for k in sd:
        if k.startswith("self."):
            assert sd[k] == ref[k], k
    assert m.has_dyn and m.is_sparse and not hasattr(m, "atlas_mask") and m.args.learn_loop_mask is False
    with pytest.raises(Exception, match="culled already"):
        m.sparsify_faces()
    # ... and it is exactly what the stage-2 model loads (MPV.py:235-262)
    from videoloop3d_b200 import MPMeshVid, default_args
    args2 = default_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpv_frm_num=3, mpi_h_scale=1.0, mpi_w_scale=1.0)
    m2 = MPMeshVid(args2, H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32),
                   1.0, 10.0)
    m2.init_from_mpi(sd)
    assert tuple(m2.atlas_dyn.shape) == (3,) + tuple(ref["atlas_dyn"].shape[1:]) and torch.equal(m2.faces_dyn, sd["faces_dyn"])
    assert MPMesh._tile_grid(274 - 77) == (9, 22, 1) and MPMesh._tile_grid(77) == (6, 13, 1)
    # a fresh stage-1 model resumes from that checkpoint (MPI.py:173-205) ...
    m3 = MPMesh(default_args_stage1(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpi_h_scale=1.0, mpi_w_scale=1.0),
                H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32), 1.0, 10.0)
    m3.init_from_mpi({k: (v.clone() if torch.is_tensor(v) else v) for k, v in ref.items()})
    sd3 = m3.state_dict()
    assert set(sd3) == set(ref) and m3.has_dyn and not hasattr(m3, "atlas_mask")
    for k, v in ref.items():
        assert (torch.equal(sd3[k], v.to(sd3[k].dtype)) if torch.is_tensor(v) else sd3[k] == v), k
    # ... and the trainer's per-step hooks exist (train_3d.py:298-301)
    m3.update_step(1000)
    assert [n for n, _ in m3.get_lrate(0)] == ["lr", "vertlr"] and abs(m3.get_lrate(100000)[0][1] - 0.1 * m3.args.lrate) < 1e-12
    with pytest.raises(NotImplementedError):
        m3.update_step(10 ** 7)


def test_tile_culling_helpers():
    """Host helpers of `MPMesh.sparsify_faces`: the atlas grid of MPI.py:367-381 (brute-force restatement) and the 3x3
    erosion / dilation with zero padding of utils.py:298-317 (naive loops)."""
    from videoloop3d_b200 import MPMesh
    for n in list(range(16, 400)) + [777, 1024, 4099, 17532]:
        h, w, filler = MPMesh._tile_grid(n)
        cands = [r for r in range(int(np.sqrt(n / 4)), int(np.sqrt(n)))]
        best = min(cands, key=lambda r: (r - n % r, cands.index(r)))          # first minimum of rows - n % rows
        assert h == best and w == n // h + 1 and filler == h * w - n and 1 <= filler <= h
        assert h * w > n and int(np.sqrt(n / 4)) <= h < int(np.sqrt(n))        # rows between sqrt(n / 4) and sqrt(n)
    assert MPMesh._tile_grid(0) == (0, 0, 0)
    with pytest.raises(ValueError):
        MPMesh._tile_grid(3)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 1, 9, 11, generator=g)
    xp = np.pad(x[0, 0].numpy(), 1)                                            # zero padding: the border erodes
    ero = np.array([[xp[r:r + 3, c:c + 3].min() for c in range(11)] for r in range(9)])
    dil = np.array([[xp[r:r + 3, c:c + 3].max() for c in range(11)] for r in range(9)])
    assert np.array_equal(MPMesh._morph(x, 1, 0)[0, 0].numpy(), ero) and np.array_equal(MPMesh._morph(x, 0, 1)[0, 0].numpy(), dil)
    ep = np.pad(ero, 1)
    ero_dil = np.array([[ep[r:r + 3, c:c + 3].max() for c in range(11)] for r in range(9)])
    assert np.array_equal(MPMesh._morph(x, 1, 1)[0, 0].numpy(), ero_dil)
    assert float(MPMesh._morph(x, 1, 0)[0, 0, 0].max()) == 0.0                 # (first row touches the zero padding)


def test_argument_validation_of_the_round2_entry_points():
    """vl3d_copy_boxes / vl3d_fused_bwd_adam_own / the sizing helpers reject bad arguments before touching the device."""
    lib = _lib.load()
    p16 = ctypes.c_void_p(16)
    assert _lib.call("vl3d_copy_boxes", None, 0, None) == 0                     # nothing to do
    with pytest.raises(_lib.Vl3dError, match="boxes"):
        _lib.call("vl3d_copy_boxes", None, _lib.MAX_BOXES + 1, None)
    with pytest.raises(_lib.Vl3dError, match="NULL"):
        _lib.call("vl3d_copy_boxes", None, 2, None)
    arr = (_lib.Box * 1)()
    arr[0].n_frames, arr[0].n_planes, arr[0].n_rows, arr[0].n_cols = 1, 1, 2, 4
    with pytest.raises(_lib.Vl3dError, match="NULL pointer"):
        _lib.call("vl3d_copy_boxes", arr, 1, None)                              # a non-empty box without pointers
    arr[0].n_rows = 0
    assert _lib.call("vl3d_copy_boxes", arr, 1, None) == 0                      # an empty box is skipped
    # sizing helpers: per-CTA scratch is two (36 x 7)-texel grids x 2 frames; the table holds a mask + 32 zones per tile
    assert lib.vl3d_fused_own_scratch_bytes(3) == 3 * 2 * 2 * 36 * 7 * 16
    tiles = ((1280 + 30) // 31) * ((720 + 6) // 7)
    assert lib.vl3d_fused_own_table_bytes(720, 1280) == 4 * (((tiles + 3) // 4) * 4 + tiles * 64)
    with pytest.raises(_lib.Vl3dError, match="own is NULL"):
        _lib.call("vl3d_fused_bwd_adam_own", None, p16, p16, None, 4, p16, p16, p16, None, p16, None, p16, p16, 1, 0.01, 0.9,
                  0.999, 6e-8, p16, 1, 2, p16, 1, p16, 0, None, p16, 0, p16, 0, None)
    ptr, size, comp = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32()
    with pytest.raises(_lib.Vl3dError, match="bad arguments"):
        _lib.call("vl3d_alloc_compressible", 0, ctypes.byref(ptr), ctypes.byref(size), ctypes.byref(comp))
    assert lib.vl3d_free_compressible(None) == 0
