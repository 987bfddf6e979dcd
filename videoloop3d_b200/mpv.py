"""`MPMeshVid` — drop-in for the reference's stage-2 model (reference: MPV.py:26-556).

Same constructor arguments, parameter / buffer names, `render` / `forward` / `lod` / `get_optimizer`
/ `get_lrate` / `update_step` / `init_from_mpi` / `state_dict` surface, so `train_3dvid.run_iter`,
`scripts/script_render_video.py` etc. can use it unchanged.  Differences that are invisible through
that surface:

  * the atlases live in HBM RGBA-interleaved (torch `channels_last`): logical shape is still
    (T,4,Hd,Wd) / (1,4,Hs,Ws), so checkpoints, optimisers and `.shape` users see the same tensors;
  * rasterise -> grid_sample -> masked_scatter -> overcompose -> smoothness (MPV.py:353-475,
    517-531) is one fused CUDA kernel (`vl3d_composite_fwd`) with a hand-written backward; the dense
    `(T,H,W,K,4)` tensor behind `variables['mpi']` is only materialised if somebody reads that key;
  * the looping loss is `vl3d_patchnn_search` + `vl3d_vote_loss`.

Nothing here imports the reference, and there is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, tiles
from ._lib import Vl3dError
from .loop_loss import Patch3DAvg, Patch3DGPNNDirectLoss, Patch3DGPNNLowMemLoss, Patch3DMSE, _check_dist
from .optim import FusedAdam


def make_depths(num_plane, min_depth, max_depth):
    """Plane depths, uniform in disparity, far -> near (reference: utils_mpi.py:210-211)."""
    return 1.0 / torch.linspace(1.0 / max_depth, 1.0 / min_depth, num_plane, dtype=torch.float32)


def get_new_intrin(old_intrin, new_h_start, new_w_start):
    """Shift the principal point for a crop starting at (h_start, w_start) (reference: utils.py:196-200)."""
    new = old_intrin.clone() if torch.is_tensor(old_intrin) else np.array(old_intrin, copy=True)
    new[..., 0, 2] -= new_w_start
    new[..., 1, 2] -= new_h_start
    return new


def gen_mpi_vertices(H, W, intrin, num_vert_h, num_vert_w, planedepth):
    """Back-project a num_vert_h x num_vert_w pixel grid to every plane (reference: utils_mpi.py:80-89)."""
    ys, xs = torch.meshgrid(torch.linspace(0, H - 1, num_vert_h), torch.linspace(0, W - 1, num_vert_w), indexing="ij")
    xy = torch.stack([xs, ys], dim=-1).reshape(1, -1, 2)
    xy = (xy - intrin[None, None, :2, 2]) * planedepth[:, None, None].type_as(xy)
    xy = xy / intrin[None, None, [0, 1], [0, 1]]
    z = planedepth[:, None, None].expand_as(xy[..., :1])
    return torch.cat([xy.reshape(-1, 2), z.reshape(-1, 1)], dim=-1)


def pose2extrin_torch(pose):
    """(.,3,4) / (.,4,4) pose <-> extrinsic (reference: utils.py:211-219)."""
    if pose.shape[-2] == 3:
        bottom = torch.zeros_like(pose[..., :1, :])
        bottom[..., 0, 3] = 1.0
        pose = torch.cat([pose, bottom], dim=-2)
    return torch.inverse(pose)


class LazyVariables(dict):
    """`variables` dict of render(): exposes the reference's keys (MPV.py:468-474); the heavy ones
    (`mpi`, `blend_weight`, `pix_to_face`) are computed by an extra kernel launch on first access."""

    _LAZY = ("mpi", "blend_weight", "pix_to_face")

    def __init__(self, eager, make_mpi):
        super().__init__(eager)
        self._make_mpi = make_mpi
        for k in self._LAZY:
            dict.__setitem__(self, k, None)
        self._done = False

    def _materialise(self):
        if not self._done:
            self._done = True
            mpi, hits = self._make_mpi()
            K = int(hits.max().item()) if hits.numel() else 0          # utils.py:64-69
            mpi = mpi[..., :K, :]
            alpha = mpi[..., -1]
            trans = torch.cumprod(1 - alpha, dim=-1)
            bw = alpha * torch.cat([torch.ones_like(alpha[..., :1]), trans[..., :-1]], dim=-1)
            dict.__setitem__(self, "mpi", mpi)
            dict.__setitem__(self, "blend_weight", bw)
            slot = torch.arange(K, device=hits.device)[None, None, :]
            dict.__setitem__(self, "pix_to_face", torch.where(slot < hits[..., None], slot, -1)[None].long())

    def __getitem__(self, k):
        if k in self._LAZY:
            self._materialise()
        return dict.__getitem__(self, k)

    # every other way of reading a value materialises too (a `None` placeholder must never leak out)
    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        self._materialise()
        return dict.items(self)

    def values(self):
        self._materialise()
        return dict.values(self)

    def copy(self):
        self._materialise()
        return dict(dict.items(self))


class MPMeshVid(nn.Module):
    """State contract (what checkpoints, optimisers and the kernels' host side rely on; reference: MPV.py:26-138):
    parameters `atlas (1,4,Hs,Ws)`, `atlas_dyn (T,4,Hd,Wd)`, `uvs`, `uvs_dyn`, `_verts`; buffers `ref_extrin`,
    `ref_intrin`, `planedepth`, `faces(_dyn)`, `uvfaces(_dyn)`.  A fresh model is dense: every quad dynamic, plane d in
    cell (d // grid_w, d % grid_w) of the dynamic atlas."""

    UNSUPPORTED = "unsupported by the vl3d kernels (and unused by every shipped stage-2 config)"

    def __init__(self, args, H, W, ref_extrin, ref_intrin, near, far):
        super().__init__()
        self._check_supported(args)
        self.args = args
        self.H, self.W, self.near, self.far = H, W, near, far
        self.frm_num, self.isloop = args.mpv_frm_num, args.mpv_isloop
        self.mpi_d, self.mpi_h_verts, self.mpi_w_verts = args.mpi_d, args.mpi_h_verts, args.mpi_w_verts
        self.rgb_mlp_type, self.use_viewdirs, self.optimize_geometry = args.rgb_mlp_type, False, False
        for k in ("swd_patch_size", "swd_patcht_size", "swd_stride", "swd_stridet"):
            setattr(self, k, getattr(args, k))
        self._init_cameras(ref_extrin, ref_intrin)
        self._init_plane_mesh(H, W, near, far)
        self._init_atlases(H, W)
        self.losses = {                                             # same keys as MPV.py:131-138 minus 'gpnn_down'
            'swd': None,                                            # (the reference's gpnn_down cannot run)
            'gpnn': Patch3DGPNNDirectLoss(),
            'gpnn_lm': Patch3DGPNNLowMemLoss(),
            'mse': Patch3DMSE,
            'avg': Patch3DAvg,
        }
        self._pack = None
        self._pack_key = None

    @classmethod
    def _check_supported(cls, args):
        if getattr(args, "fp16", False):
            raise Vl3dError("fp16 is marked 'do NOT use' in the reference (config_parser.py:32-33); fp32 only")
        if getattr(args, "atlas_cnl", 4) != 4 or getattr(args, "rgb_mlp_type", "direct") != "direct":
            raise Vl3dError(f"atlas_cnl != 4 / rgb_mlp_type != direct: {cls.UNSUPPORTED}")
        if args.rgb_activate != "sigmoid" or args.alpha_activate != "sigmoid":
            raise Vl3dError(f"non-sigmoid activations: {cls.UNSUPPORTED}")
        if args.mpi_d > 32:
            raise Vl3dError("mpi_d > 32 is not supported by the composite kernel")
        if args.mpi_d % args.atlas_grid_h:
            raise AssertionError("mpi_d and atlas_grid_h should match")

    def _init_cameras(self, ref_extrin, ref_intrin):
        ref_extrin, ref_intrin = np.asarray(ref_extrin), np.asarray(ref_intrin)
        assert ref_extrin.shape == (4, 4) and ref_intrin.shape == (3, 3)
        self.register_buffer("ref_extrin", torch.tensor(ref_extrin))
        self.register_buffer("ref_intrin", torch.tensor(ref_intrin).float())

    def _init_plane_mesh(self, H, W, near, far):
        """D fronto-parallel planes, uniform in disparity, nearest first; each a regular (hv x wv) vertex grid that spans
        the scaled image rectangle, centred on the reference view (MPV.py:47-71)."""
        args, D, hv, wv = self.args, self.mpi_d, self.mpi_h_verts, self.mpi_w_verts
        self._mpi_hw = (int(args.mpi_h_scale * H), int(args.mpi_w_scale * W))
        self.H_start, self.W_start = (self._mpi_hw[0] - H) // 2, (self._mpi_hw[1] - W) // 2
        self.register_buffer("planedepth", make_depths(D, near, far).float().flip(0))
        centred = get_new_intrin(self.ref_intrin, -self.H_start, -self.W_start)
        verts = gen_mpi_vertices(*self._mpi_hw, centred, hv, wv, self.planedepth)
        if args.normalize_verts:
            verts = (verts.reshape(D, -1) / self.planedepth[:, None]).reshape_as(verts)
        self._verts = nn.Parameter(verts, requires_grad=True)
        quads = torch.from_numpy(tiles.quad_grid_faces(D, hv, wv))
        self.register_buffer("faces", quads[:0].clone())           # no static tiles until init_from_mpi
        self.register_buffer("faces_dyn", quads)
        self.register_buffer("uvfaces", quads[:0].clone())
        self.register_buffer("uvfaces_dyn", quads.clone())          # dense layout: uv vertices == mesh vertices

    def _init_atlases(self, H, W):
        """Dense layout: the planes tile the dynamic atlas on an atlas_grid_h x (D / atlas_grid_h) grid (MPV.py:37-44,
        75-104); texels N(0, init_std), alpha logits -2.  The static atlas has the same size and is unused until a
        stage-1 result is loaded.  RNG order (static first) as in the reference, so seeded runs start identically."""
        args = self.args
        gh, gw = args.atlas_grid_h, self.mpi_d // args.atlas_grid_h
        self.is_sparse = self.has_dyn = False
        self.atlas_grid_dyn_h, self.atlas_grid_dyn_w = self.atlas_grid_h, self.atlas_grid_w = gh, gw
        full = (int(gh * self._mpi_hw[0]), int(gw * self._mpi_hw[1]))
        self.atlas_full_dyn_h, self.atlas_full_dyn_w = self.atlas_full_h, self.atlas_full_w = full
        uv = tiles.dense_atlas_uvs(gh, gw, self.mpi_h_verts, self.mpi_w_verts)
        self.register_parameter("uvs", nn.Parameter(uv[:0].clone(), requires_grad=True))
        self.register_parameter("uvs_dyn", nn.Parameter(uv, requires_grad=True))
        atlas = torch.rand((1, 4) + full)
        atlas_dyn = torch.randn((self.frm_num, 4) + full) * args.init_std
        atlas[:, -1] = -2                                                          # MPV.py:109-110
        atlas_dyn[:, -1] = -2
        self.register_parameter("atlas_dyn", nn.Parameter(ops.as_texels(atlas_dyn), requires_grad=True))
        self.register_parameter("atlas", nn.Parameter(ops.as_texels(atlas), requires_grad=True))

    # ------------------------------------------------------------------ geometry cache
    @property
    def verts(self):
        verts = self._verts
        if self.args.normalize_verts:
            verts = (verts.reshape(len(self.planedepth), -1) * self.planedepth[:, None]).reshape_as(verts)
        return verts

    def invalidate_geometry(self):
        self._pack = None
        self._ref_inv = None

    def ref_extrin_inv_host(self):
        """inverse of `ref_extrin` as a float64 host array, cached until the geometry is invalidated (the per-step
        `extrin = tar_extrin @ ref_extrin^-1` of MPV.py:478 then needs no device round trip)."""
        key = (self.ref_extrin.data_ptr(), self.ref_extrin._version)
        if getattr(self, "_ref_inv", None) is None or self._ref_inv[0] != key:
            self._ref_inv = (key, np.linalg.inv(self.ref_extrin.detach().double().cpu().numpy()))
        return self._ref_inv[1]

    def _texels(self):
        """Keep both atlases in the RGBA-interleaved layout (re-home them once if something replaced
        `.data` with a differently-strided tensor, e.g. init_from_mpi's stride-0 expand, MPV.py:260-262)."""
        for name in ("atlas", "atlas_dyn"):
            p = getattr(self, name)
            t = ops.as_texels(p.data)
            if t is not p.data:
                p.data = t
        return self.atlas_dyn, self.atlas

    def mesh_pack(self):
        key = (self.faces.data_ptr(), self.faces_dyn.data_ptr(), self.uvs.data_ptr(), self.uvs_dyn.data_ptr(),
               self._verts.data_ptr(), tuple(self.atlas.shape), tuple(self.atlas_dyn.shape), self.uvs._version,
               self.uvs_dyn._version, self._verts._version, str(self.atlas_dyn.device))
        if self._pack is None or key != self._pack_key:
            if not self.atlas_dyn.is_cuda:
                raise Vl3dError("MPMeshVid must be moved to a CUDA device before rendering (no CPU fallback)")
            self._pack = ops.make_mesh_pack(
                dict(verts=self.verts, faces=self.faces, uvs=self.uvs, uvfaces=self.uvfaces,
                     atlas_hw=tuple(self.atlas.shape[-2:]), faces_dyn=self.faces_dyn, uvs_dyn=self.uvs_dyn,
                     uvfaces_dyn=self.uvfaces_dyn, atlas_dyn_hw=tuple(self.atlas_dyn.shape[-2:])),
                self.mpi_d, self.mpi_h_verts, self.mpi_w_verts, self.atlas_dyn.device)
            self._pack_key = key
        return self._pack

    def make_view(self, H, W, extrin, intrin):
        """extrin: ref->target (already multiplied by ref_extrin^-1, as `render` receives it)."""
        if torch.is_tensor(extrin) and extrin.dim() == 3:
            assert extrin.shape[0] == 1, "batching of views is not supported (MPV.py:388)"
        return ops.make_view(self.mesh_pack(), H, W, extrin, intrin, np.eye(4),
                             tuple(self.atlas_dyn.shape[-2:]), tuple(self.atlas.shape[-2:]))

    def _ts_tensor(self, ts):
        ts = torch.as_tensor(np.asarray(ts.cpu() if torch.is_tensor(ts) else ts)).reshape(-1)
        if len(ts) == 0:
            raise ValueError("ts is empty")
        if int(ts.min()) < 0 or int(ts.max()) >= self.atlas_dyn.shape[0]:
            raise IndexError(f"frame index out of range [0, {self.atlas_dyn.shape[0]})")
        return ts.to(device=self.atlas_dyn.device, dtype=torch.int32)

    # ------------------------------------------------------------------ render / forward
    def _render_planar(self, H, W, extrin, intrin, ts, pad=0, smooth=False):
        atlas_dyn, atlas = self._texels()
        view = self.make_view(H, W, extrin, intrin)
        identity = ts is None
        ts_t = None if identity else self._ts_tensor(ts)
        T = atlas_dyn.shape[0] if identity else len(ts_t)
        rgb, alpha, sums = ops.CompositeFn.apply(atlas_dyn, atlas, view, self._pack, ts_t, T, pad, smooth)
        return rgb, alpha, sums, view, ts_t, T

    def _bg_color(self):
        """None, or the (3,) background of MPV.py:455-460 ('random' draws from torch's CPU generator like the reference)."""
        bg = getattr(self.args, "bg_color", "")
        if len(bg) == 0:
            return None
        if bg == "random":
            return torch.rand(3)
        return torch.tensor([float(v) for v in bg.split('#')], dtype=torch.float32)

    def _render_terms(self, view, ts_t, T, H, W, extrin, intrin, want_disp=False, want_sparsity=False):
        """Differentiable alpha (T,H,W), disp (T,H,W) and sparsity sum of the same view (csrc/terms.cu): the optional
        terms of MPV.py:454-466,511-515 that every shipped stage-2 config leaves off."""
        atlas_dyn, atlas = self._texels()
        inv_depth = ops.make_inv_depth(self._pack, H, W, extrin, intrin, np.eye(4)) if want_disp else None
        return ops.CompositeTermsFn.apply(atlas_dyn, atlas, view, self._pack, ts_t, T, inv_depth, 1e-4, want_disp,
                                          want_sparsity)

    @staticmethod
    def _blend_bg(rgb, alpha, bg):
        """MPV.py:461 on the planar layout: rgb (T,3,H,W), alpha (T,H,W)."""
        a = alpha[:, None]
        return rgb * a + bg.to(rgb)[None, :, None, None] * (-a + 1)

    def render(self, H, W, extrin, intrin, ts):
        """rgb (len(ts),H,W,3), variables  (reference: MPV.py:351-475)."""
        rgb, alpha, _, view, ts_t, T = self._render_planar(H, W, extrin, intrin, ts)
        bg = self._bg_color()
        want_disp = getattr(self.args, "d_smooth_loss_weight", 0) > 0
        disp = None
        if bg is not None or want_disp:
            alpha, disp_t, _ = self._render_terms(view, ts_t, T, H, W, extrin, intrin, want_disp=want_disp)
            if bg is not None:
                rgb = self._blend_bg(rgb, alpha, bg)
            if want_disp:
                disp = disp_t

        def make_mpi():
            with torch.no_grad():
                _, _, mpi, hits = ops.composite_fwd(view, self._pack, self.atlas_dyn, self.atlas, ts_t, T, 0,
                                                    want_mpi=True, want_hits=True)
            return mpi, hits

        variables = LazyVariables({"disp_norm": disp, "alpha": alpha}, make_mpi)
        return rgb.permute(0, 2, 3, 1), variables

    def forward(self, h, w, tar_extrins, tar_intrins, ts=None, res=None, losscfg=None):
        """Train: (None, {swd, rgb_smooth, a_smooth[, sparsity, density, d_smooth]}) each (1,1); eval: (rgb (T,3,H,W), {})
        (reference: MPV.py:477-556)."""
        # MPV.py:478, evaluated in float64 on the host (where the view descriptor is built)
        tar = tar_extrins.detach().double().cpu().numpy() if torch.is_tensor(tar_extrins) else np.asarray(tar_extrins, np.float64)
        extrins = tar.reshape(-1, 4, 4)[:1] @ self.ref_extrin_inv_host()
        bg = self._bg_color()
        if not self.training:
            rgb, _, _, view, ts_t, T = self._render_planar(h, w, extrins, tar_intrins, ts)
            if bg is not None:
                alpha, _, _ = self._render_terms(view, ts_t, T, h, w, extrins, tar_intrins)
                rgb = self._blend_bg(rgb, alpha, bg)
            return rgb, {}

        assert res is not None
        args = self.args
        want_sp = getattr(args, "sparsity_loss_weight", 0) > 0
        want_den = getattr(args, "density_loss_weight", 0) > 0
        want_ds = getattr(args, "d_smooth_loss_weight", 0) > 0
        cfg = {k: (v[0].item() if torch.is_tensor(v) else v[0]) for k, v in losscfg.items()}   # MPV.py:494
        _check_dist(cfg)
        loss_name = cfg.pop('loss_name')
        loss_gain = float(cfg.pop('loss_gain', 1.))
        loss = self.losses[loss_name]
        smooth = args.rgb_smooth_loss_weight > 0 or args.a_smooth_loss_weight > 0
        pad = self.swd_patcht_size - 1 if self.isloop else 0                       # MPV.py:490-492
        # the loop pad is written by the render kernel unless a background blend sits between render and pad
        rgb_pad, _, sums, view, ts_t, T = self._render_planar(h, w, extrins, tar_intrins, ts, pad=0 if bg is not None else pad,
                                                              smooth=smooth)
        alpha = disp = sp_sum = None
        if bg is not None or want_sp or want_den or want_ds:
            alpha, disp, sp_sum = self._render_terms(view, ts_t, T, h, w, extrins, tar_intrins, want_disp=want_ds,
                                                     want_sparsity=want_sp)
        if bg is not None:                                                         # MPV.py:455-461, then MPV.py:490-492
            rgb = self._blend_bg(rgb_pad, alpha, bg)
            rgb_pad = torch.cat([rgb, rgb[:pad]], 0) if pad else rgb.contiguous()
        res0 = res[0]
        if not res0.is_contiguous():
            res0 = res0.contiguous()
        xscale = None
        if args.scale_invariant:                                                   # MPV.py:499-504
            xscale = ops.scale_invariant(rgb_pad.detach(), T, res0)
        if isinstance(loss, (Patch3DGPNNLowMemLoss, Patch3DGPNNDirectLoss)):
            _ = cfg.pop("macro_block", None), cfg.pop("factor", None), cfg.pop("dist_fn", None)
            main_loss = loss.planar(rgb_pad, xscale, res0, cfg)
        else:
            x = rgb_pad if xscale is None else rgb_pad * xscale
            main_loss = loss(x.permute(1, 0, 2, 3)[None], res.permute(0, 2, 1, 3, 4), **cfg)
        extra = {'swd': main_loss.reshape(1, -1) * loss_gain}
        if want_sp:                                                                # MPV.py:511-515
            val = sp_sum / (T * h * w) / np.sqrt(self.mpi_d) * loss_gain
            extra["sparsity"] = val.float().reshape(1, -1)
        # slot-wise smoothness: mean|dx| + mean|dy| over (T,H,W,K,c), times gain*K/D: K cancels (MPV.py:517-531)
        nx = max(T * h * (w - 1), 1) * self.mpi_d
        ny = max(T * (h - 1) * w, 1) * self.mpi_d
        if args.rgb_smooth_loss_weight > 0:
            val = (sums[0] / (3 * nx) + sums[1] / (3 * ny)) * loss_gain
            extra["rgb_smooth"] = val.float().reshape(1, -1)
        if args.a_smooth_loss_weight > 0:
            val = (sums[2] / nx + sums[3] / ny) * loss_gain
            extra["a_smooth"] = val.float().reshape(1, -1)
        if want_den:                                                               # MPV.py:533-536
            extra["density"] = (alpha - 1).abs().mean().reshape(1, -1)
        if want_ds:                                                                # MPV.py:538-551
            gx = (disp[:, 1:, :-1] - disp[:, 1:, 1:]).abs()
            gy = (disp[:, :-1, 1:] - disp[:, 1:, 1:]).abs()
            extra["d_smooth"] = (gx + gy).mean().reshape(1, -1)
        return None, extra

    # ------------------------------------------------------------------ level of detail (MPV.py:140-198)
    @torch.no_grad()
    def lod(self, factor):
        resize = lambda a, hw: F.interpolate(a, size=hw, mode="bilinear", align_corners=False, antialias=False)
        if not self.is_sparse:
            h, w = int(self.atlas_full_dyn_h * factor), int(self.atlas_full_dyn_w * factor)
            new_atlas = resize(self.atlas_dyn.data.contiguous(), (h, w))
            self.register_parameter("atlas_dyn", nn.Parameter(ops.as_texels(new_atlas), requires_grad=True))
        else:
            atlas_h, atlas_w = self.atlas.shape[-2:]
            gridh, gridw = self.atlas_grid_h, self.atlas_grid_w
            tileh, tilew = atlas_h // gridh, atlas_w // gridw
            fullh, fullw = self.atlas_full_h // gridh, self.atlas_full_w // gridw
            newh, neww = max(int(fullh * factor), 2), max(int(fullw * factor), 2)

            def resize_tiles(a, gh, gw):
                b, c = a.shape[:2]
                a = a.reshape(b, c, gh, tileh, gw, tilew).permute(0, 2, 4, 1, 3, 5).reshape(-1, c, tileh, tilew)
                a = resize(a.contiguous(), (newh, neww))
                a = a.reshape(b, gh, gw, c, newh, neww).permute(0, 3, 1, 4, 2, 5)
                return a.reshape(b, c, gh * newh, gw * neww)

            def realign(uvs, old_h, old_w, new_h, new_w):
                out = []
                for col, old, new, tile, newtile in ((0, old_w, new_w, tilew, neww), (1, old_h, new_h, tileh, newh)):
                    pix = torch.round((uvs[:, col] + 1) / 2 * (old - 1)).long()
                    tile_idx, tile_pix = pix // tile, pix % tile
                    assert torch.all((tile_pix == 0) | (tile_pix == tile - 1))
                    tile_pix = torch.where(tile_pix == tile - 1, torch.full_like(tile_pix, newtile - 1), tile_pix)
                    out.append((tile_idx * newtile + tile_pix) / (new - 1) * 2 - 1)
                return torch.stack(out, dim=1).to(uvs.dtype)

            new_atlas = resize_tiles(self.atlas.data.contiguous(), gridh, gridw)
            self.register_parameter("atlas", nn.Parameter(ops.as_texels(new_atlas), requires_grad=True))
            self.uvs.data = realign(self.uvs.data, atlas_h, atlas_w, *self.atlas.shape[-2:])
            if self.has_dyn:
                dh, dw = self.atlas_dyn.shape[-2:]
                new_dyn = resize_tiles(self.atlas_dyn.data.contiguous(), self.atlas_grid_dyn_h, self.atlas_grid_dyn_w)
                self.register_parameter("atlas_dyn", nn.Parameter(ops.as_texels(new_dyn), requires_grad=True))
                self.uvs_dyn.data = realign(self.uvs_dyn.data, dh, dw, *self.atlas_dyn.shape[-2:])
        self.invalidate_geometry()

    # ------------------------------------------------------------------ optimiser (MPV.py:200-233)
    def get_optimizer(self, step):
        args = self.args
        (_, base_lr), (_, verts_lr) = self.get_lrate(step)
        named = dict(self.named_parameters())
        base = [v for k, v in named.items() if k != "_verts"]
        params = [{'params': base}, {'params': [named["_verts"]], 'lr': verts_lr}]
        if args.optimizer == 'adam':
            return FusedAdam(params, lr=base_lr, betas=(0.9, 0.999), eps=6e-8)
        if args.optimizer == 'sgd':
            return torch.optim.SGD(params=params, lr=base_lr, momentum=0.9)
        raise RuntimeError(f"Unrecongnized optimizer type {args.optimizer}")

    def get_lrate(self, step):
        args = self.args
        scaling = 0.1 ** (step / (args.lrate_decay * 1000))
        return [("lr", args.lrate * scaling), ("vertlr", args.lrate * args.optimize_verts_gain * scaling)]

    def update_step(self, step):
        if step >= self.args.optimize_geo_start:
            raise NotImplementedError("geometry optimisation (optimize_geo_start) is never reached by the "
                                      "reference's configs and is not supported")

    # ------------------------------------------------------------------ checkpoints (MPV.py:235-304)
    _SCALARS = ("is_sparse", "atlas_full_w", "atlas_full_h", "atlas_grid_h", "atlas_grid_w")
    _SCALARS_DYN = ("has_dyn", "atlas_full_dyn_w", "atlas_full_dyn_h", "atlas_grid_dyn_h", "atlas_grid_dyn_w")

    def init_from_mpi(self, state_dict):
        sd = state_dict
        self._verts.data = sd['_verts'].type_as(self._verts)
        self.ref_extrin.data = sd['ref_extrin'].type_as(self.ref_extrin)
        self.ref_intrin.data = sd['ref_intrin'].type_as(self.ref_intrin)
        self.planedepth.data = sd['planedepth'].type_as(self.planedepth)
        n_frames = len(self.atlas_dyn)
        dev = self.atlas_dyn.device

        def texels(a, frames=None):
            a = a.to(device=dev, dtype=torch.float32)
            if frames is not None:
                a = a.expand(frames, -1, -1, -1)        # one stage-1 frame replicated over T (MPV.py:260-262)
            return ops.as_texels(a)

        if "self.has_dyn" in sd.keys():
            self.uvs.data = sd['uvs'].type_as(self.uvs)
            self.atlas.data = texels(sd['atlas'])
            self.uvfaces.data = sd['uvfaces'].type_as(self.uvfaces)
            self.faces.data = sd['faces'].type_as(self.faces)
            for k in self._SCALARS + self._SCALARS_DYN:
                setattr(self, k, sd["self." + k])
            self.uvs_dyn.data = sd['uvs_dyn'].type_as(self.uvs)
            self.uvfaces_dyn.data = sd['uvfaces_dyn'].type_as(self.uvfaces)
            self.faces_dyn.data = sd['faces_dyn'].type_as(self.faces)
            src = sd['atlas_dyn']
            self.atlas_dyn.data = texels(src, n_frames if src.shape[0] == 1 else None)
            if self.frm_num != len(self.atlas_dyn):
                self.frm_num = len(self.atlas_dyn)
        else:                                           # static MPI loaded as the dynamic part (MPV.py:264-288)
            self.uvs.data = sd['uvs'][:0].clone().type_as(self.uvs)
            self.atlas.data = texels(sd['atlas'][:, :, :1, :1])
            self.uvfaces.data = sd['uvfaces'][:0].clone().type_as(self.uvfaces)
            self.faces.data = sd['faces'][:0].clone().type_as(self.faces)
            for k in self._SCALARS:
                setattr(self, k, sd["self." + k])
            self.atlas_full_dyn_w, self.atlas_full_dyn_h = sd["self.atlas_full_w"], sd["self.atlas_full_h"]
            self.atlas_grid_dyn_h, self.atlas_grid_dyn_w = sd["self.atlas_grid_h"], sd["self.atlas_grid_w"]
            self.uvs_dyn.data = sd['uvs'].type_as(self.uvs)
            self.uvfaces_dyn.data = sd['uvfaces'].type_as(self.uvfaces)
            self.faces_dyn.data = sd['faces'].type_as(self.faces)
            self.atlas_dyn.data = texels(sd['atlas'], n_frames)
        self.invalidate_geometry()

    def state_dict(self, destination=None, prefix='', keep_vars=False):
        sd = super().state_dict()
        for k in self._SCALARS:
            sd["self." + k] = getattr(self, k)
        if hasattr(self, "atlas_dyn"):
            for k in self._SCALARS_DYN:
                sd["self." + k] = getattr(self, k)
        return sd
