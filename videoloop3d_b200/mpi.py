"""`MPMesh` — drop-in for the `render` / `forward` surface of the reference's stage-1 model (reference: MPI.py:38-124,
452-652; SURVEY.md §8(f) N4, second half).

Stage 1 fits ONE static multiplane image (plus a one-channel loop-mask atlas) to the averaged input views before
stage 2 turns it into the looping video; it is the same composite as stage 2 with one frame: every quad samples the
static atlas.  It is built from the stage-2 kernels, nothing new on the device:

  * rgb / smoothness sums            `vl3d_composite_fwd / _bwd` (static tiles, T = 1),
  * alpha, disparity, sparsity       `vl3d_composite_terms_fwd / _bwd` (csrc/terms.cu); MPI.py:553's normalisation of the
                                     inverse depth is affine and is folded into the per-plane coefficients on the host,
  * loop-mask label (MPI.py:568-580) a second composite over the texel tensor (mask, mask, mask, alpha.detach()): its
                                     "rgb" is the label, its smoothness sums are 3x the l_smooth sums, and the detach
                                     keeps the geometry independent of the mask exactly as the reference does.

Supported: `rgb_mlp_type='direct'`, sigmoid activations, one view per call (the reference's own batching of views does
not run: MPI.py:472 broadcasts (B,3,3) against (1,N,3,1)).  `sparsify_faces` (tile culling, MPI.py:289-442) and the
checkpoint format (`state_dict`, MPI.py:207-221) are here too, so a stage-1 result flows into `MPMeshVid.init_from_mpi`; after
culling the model carries static AND one-frame dynamic tiles and keeps rendering through the same kernels.  Not built: the
stage-1 trainer's data side (`train_3d.py`: image loading, loopable-mask estimation) and the mesh / texture exporters.
There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, tiles
from ._lib import Vl3dError
from .mpv import LazyVariables, gen_mpi_vertices, get_new_intrin, make_depths

ALPHA_INIT_VAL = -3.0                                               # MPI.py:35


class Stage1Variables(LazyVariables):
    """`variables` of MPMesh.render (MPI.py:585-592): `loopmask3d` joins the lazily materialised keys and
    `blend_weight` honours `normalize_blendweight_fordepth` (MPI.py:564-565)."""

    _LAZY = LazyVariables._LAZY + ("loopmask3d",)

    def __init__(self, eager, make_mpi, make_mask_mpi=None, normalize_bw=False):
        super().__init__(eager, make_mpi)
        self._make_mask_mpi, self._normalize_bw = make_mask_mpi, normalize_bw

    def _materialise(self):
        if self._done:
            return
        super()._materialise()
        K = dict.__getitem__(self, "mpi").shape[-2]
        if self._normalize_bw:
            bw = dict.__getitem__(self, "blend_weight")
            dict.__setitem__(self, "blend_weight", bw / dict.__getitem__(self, "alpha").detach().clamp_min(1e-10)[..., None])
        if self._make_mask_mpi is not None:
            dict.__setitem__(self, "loopmask3d", self._make_mask_mpi()[..., :K, :1])


class MPMesh(nn.Module):
    """State contract (reference: MPI.py:38-124): parameters `atlas (1,4,Ha,Wa)`, `atlas_mask (1,1,Ha,Wa)` (with
    `learn_loop_mask`), `uvs`, `_verts`; buffers `ref_extrin`, `ref_intrin`, `planedepth`, `faces`, `uvfaces`.  A fresh
    model is dense: plane d occupies cell (d // grid_w, d % grid_w) of the atlas."""

    def __init__(self, args, H, W, ref_extrin, ref_intrin, near, far):
        super().__init__()
        if getattr(args, "rgb_mlp_type", "direct") != "direct":
            raise Vl3dError("rgb_mlp_type != 'direct' (view-dependent spherical harmonics) is not supported by the vl3d kernels")
        if args.rgb_activate != "sigmoid" or args.alpha_activate != "sigmoid":
            raise Vl3dError("non-sigmoid activations are not supported by the vl3d kernels")
        if args.mpi_d > 32:
            raise Vl3dError("mpi_d > 32 is not supported by the composite kernel")
        assert args.mpi_d % args.atlas_grid_h == 0, "mpi_d and atlas_grid_h should match"
        self.args = args
        self.upsample_stage = getattr(args, "upsample_stage", "")
        D, hv, wv = args.mpi_d, args.mpi_h_verts, args.mpi_w_verts
        mpi_h, mpi_w = int(args.mpi_h_scale * H), int(args.mpi_w_scale * W)
        self.mpi_d, self.near, self.far = D, near, far
        self.mpi_h_verts, self.mpi_w_verts, self.mpi_h, self.mpi_w, self.H, self.W = hv, wv, mpi_h, mpi_w, H, W
        self.atlas_grid_h, self.atlas_grid_w = args.atlas_grid_h, D // args.atlas_grid_h
        self.is_sparse = self.has_dyn = False
        self.atlas_full_h, self.atlas_full_w = int(self.atlas_grid_h * mpi_h), int(self.atlas_grid_w * mpi_w)
        ref_extrin, ref_intrin = np.asarray(ref_extrin), np.asarray(ref_intrin)
        assert ref_extrin.shape == (4, 4) and ref_intrin.shape == (3, 3)
        self.register_buffer("ref_extrin", torch.tensor(ref_extrin))
        self.register_buffer("ref_intrin", torch.tensor(ref_intrin).float())
        self.register_buffer("planedepth", make_depths(D, near, far).float().flip(0))
        self.H_start, self.W_start = (mpi_h - H) // 2, (mpi_w - W) // 2
        verts = gen_mpi_vertices(mpi_h, mpi_w, get_new_intrin(self.ref_intrin, -self.H_start, -self.W_start), hv, wv,
                                 self.planedepth)
        if args.normalize_verts:
            verts = (verts.reshape(D, -1) / self.planedepth[:, None]).reshape_as(verts)
        quads = torch.from_numpy(tiles.quad_grid_faces(D, hv, wv))
        self.register_buffer("uvfaces", quads.clone())
        self._verts = nn.Parameter(verts, requires_grad=True)
        self.register_buffer("faces", quads)
        self.optimize_geometry = False
        uv = tiles.dense_atlas_uvs(self.atlas_grid_h, self.atlas_grid_w, hv, wv)
        self.register_parameter("uvs", nn.Parameter(uv, requires_grad=True))
        self.rgb_mlp_type, self.use_viewdirs = "direct", False
        atlas = torch.rand((1, 4, self.atlas_full_h, self.atlas_full_w))       # the constructor's only RNG draw (MPI.py:102)
        atlas[:, -1] = ALPHA_INIT_VAL
        self.register_parameter("atlas", nn.Parameter(ops.as_texels(atlas), requires_grad=True))
        if args.learn_loop_mask:
            self.register_parameter("atlas_mask", nn.Parameter(torch.ones_like(atlas[:, :1]) * ALPHA_INIT_VAL, requires_grad=True))
        self._pack = self._pack_key = self._ref_inv = self._no_dyn = None

    # ------------------------------------------------------------------ geometry cache
    @property
    def verts(self):
        verts = self._verts
        if self.args.normalize_verts:
            verts = (verts.reshape(len(self.planedepth), -1) * self.planedepth[:, None]).reshape_as(verts)
        return verts

    def invalidate_geometry(self):
        self._pack = self._ref_inv = None

    def ref_extrin_inv_host(self):
        key = (self.ref_extrin.data_ptr(), self.ref_extrin._version)
        if self._ref_inv is None or self._ref_inv[0] != key:
            self._ref_inv = (key, np.linalg.inv(self.ref_extrin.detach().double().cpu().numpy()))
        return self._ref_inv[1]

    def _texels(self):
        """(atlas_dyn, atlas) in the RGBA-interleaved layout; before culling the dynamic atlas is a 1x1 dummy."""
        for name in ("atlas", "atlas_dyn"):
            p = getattr(self, name, None)
            if p is not None:
                t = ops.as_texels(p.data)
                if t is not p.data:
                    p.data = t
        if self.has_dyn:
            return self.atlas_dyn, self.atlas
        if getattr(self, "_no_dyn", None) is None or self._no_dyn.device != self.atlas.device:
            self._no_dyn = torch.zeros((1, 4, 1, 1), dtype=torch.float32, device=self.atlas.device)
        return self._no_dyn, self.atlas

    def mesh_pack(self):
        dyn = self.has_dyn
        key = (self.faces.data_ptr(), self.uvs.data_ptr(), self._verts.data_ptr(), tuple(self.atlas.shape), self.uvs._version,
               self._verts._version, str(self.atlas.device), dyn, self.faces_dyn.data_ptr() if dyn else 0,
               tuple(self.atlas_dyn.shape) if dyn else None)
        if self._pack is None or key != self._pack_key:
            if not self.atlas.is_cuda:
                raise Vl3dError("MPMesh must be moved to a CUDA device before rendering (no CPU fallback)")
            none3 = torch.zeros(0, 3, dtype=torch.long)
            self._pack = ops.make_mesh_pack(
                dict(verts=self.verts, faces=self.faces, uvs=self.uvs, uvfaces=self.uvfaces, atlas_hw=tuple(self.atlas.shape[-2:]),
                     faces_dyn=self.faces_dyn if dyn else none3, uvs_dyn=self.uvs_dyn if dyn else torch.zeros(0, 2),
                     uvfaces_dyn=self.uvfaces_dyn if dyn else none3,
                     atlas_dyn_hw=tuple(self.atlas_dyn.shape[-2:]) if dyn else (1, 1)),
                self.mpi_d, self.mpi_h_verts, self.mpi_w_verts, self.atlas.device)
            self._pack_key = key
        return self._pack

    def _bg_color(self):
        bg = getattr(self.args, "bg_color", "")
        if len(bg) == 0:
            return None
        if bg == "random":
            return torch.rand(3)                                    # CPU generator, like the reference (MPI.py:555)
        return torch.tensor([float(v) for v in bg.split('#')], dtype=torch.float32)

    # ------------------------------------------------------------------ render / forward
    def render(self, H, W, extrin, intrin):
        """rgbl (1,H,W,3 or 4), variables  (reference: MPI.py:452-594).  `extrin`: ref -> target, (1,4,4)."""
        if len(extrin) != 1:
            raise Vl3dError("MPMesh.render takes one view per call (the reference's batching does not run either, MPI.py:472)")
        args = self.args
        if self.has_dyn and args.learn_loop_mask:
            raise AssertionError("learn_loop_mask with dynamic tiles (MPI.py:569 asserts the same)")
        if getattr(args, "add_uv_noise", False) and self.training:
            raise NotImplementedError("add_uv_noise is off in every shipped config and not supported")
        pack = self.mesh_pack()
        atlas_dyn, atlas = self._texels()
        view = ops.make_view(pack, H, W, extrin, intrin, np.eye(4), tuple(atlas_dyn.shape[-2:]), tuple(atlas.shape[-2:]))
        train = self.training
        w_of = lambda k: getattr(args, f"{k}_loss_weight", 0) if train else 0
        smooth = w_of("rgb_smooth") > 0 or w_of("a_smooth") > 0
        rgb, _, sums = ops.CompositeFn.apply(atlas_dyn, atlas, view, pack, None, 1, 0, smooth)
        span = 1.0 / self.near - 1.0 / self.far                     # MPI.py:552-553: (1/z - 1/far) / (1/near - 1/far)
        inv_depth = ops.make_inv_depth(pack, H, W, extrin, intrin, np.eye(4), scale=1.0 / span, offset=-1.0 / (self.far * span))
        alpha, disp, sp_sum = ops.CompositeTermsFn.apply(atlas_dyn, atlas, view, pack, None, 1, inv_depth, 1e-6, True,
                                                         w_of("sparsity") > 0)
        bg = self._bg_color()
        if bg is not None:                                          # MPI.py:554-560
            a = alpha[:, None]
            rgb = rgb * a + bg.to(rgb)[None, :, None, None] * (-a + 1)
        normalize_bw = bool(getattr(args, "normalize_blendweight_fordepth", False))
        if normalize_bw:                                            # MPI.py:563-566
            disp = disp / alpha.clamp_min(1e-10)
        rgbl = rgb.permute(0, 2, 3, 1)
        make_mask_mpi, lsums = None, None
        if args.learn_loop_mask:                                    # MPI.py:568-580
            mask_texels = self._mask_texels(atlas)
            lab, _, lsums = ops.CompositeFn.apply(self._no_dyn, mask_texels, view, pack, None, 1, 0, w_of("l_smooth") > 0)
            rgbl = torch.cat([rgbl, lab[:, :1].permute(0, 2, 3, 1)], dim=-1)

            def make_mask_mpi():
                with torch.no_grad():
                    return ops.composite_fwd(view, pack, self._no_dyn, mask_texels.detach(), None, 1, 0, want_mpi=True)[2]

        def make_mpi():
            with torch.no_grad():
                _, _, mpi, hits = ops.composite_fwd(view, pack, atlas_dyn.detach(), atlas.detach(), None, 1, 0, want_mpi=True, want_hits=True)
            return mpi, hits

        variables = Stage1Variables({"disp_norm": disp, "alpha": alpha}, make_mpi, make_mask_mpi, normalize_bw)
        variables.train_sums = dict(smooth=sums, sparsity=sp_sum, l_smooth=lsums)   # consumed by forward()
        return rgbl, variables

    def _mask_texels(self, atlas):
        """(mask, mask, mask, alpha.detach()) as an RGBA-interleaved texel tensor, differentiable w.r.t. `atlas_mask`."""
        fake = torch.cat([self.atlas_mask.expand(-1, 3, -1, -1), atlas[:, 3:4].detach()], dim=1)
        return ops.as_texels(fake.contiguous(memory_format=torch.channels_last))

    def forward(self, h, w, tar_extrins, tar_intrins):
        """(rgbl (1,C,h,w), extra) with extra = {sparsity, rgb_smooth, a_smooth, d_smooth, l_smooth, density}, each (1,1), in
        training mode and {} otherwise (reference: MPI.py:596-652)."""
        tar = tar_extrins.detach().double().cpu().numpy() if torch.is_tensor(tar_extrins) else np.asarray(tar_extrins, np.float64)
        extrins = tar.reshape(-1, 4, 4) @ self.ref_extrin_inv_host()
        rgbl, variables = self.render(h, w, extrins, tar_intrins)
        rgbl = rgbl.permute(0, 3, 1, 2)
        extra = {}
        if not self.training:
            return rgbl, extra
        args, D = self.args, self.mpi_d
        ts = variables.train_sums
        nx, ny = max(h * (w - 1), 1) * D, max((h - 1) * w, 1) * D    # mean over (1,H,W-1,K,c) times K/D: K cancels
        if args.sparsity_loss_weight > 0:                           # MPI.py:603-607
            extra["sparsity"] = (ts["sparsity"] / (h * w) / np.sqrt(D)).float().reshape(1, -1)
        if args.rgb_smooth_loss_weight > 0:                         # MPI.py:609-615
            extra["rgb_smooth"] = (ts["smooth"][0] / (3 * nx) + ts["smooth"][1] / (3 * ny)).float().reshape(1, -1)
        if args.a_smooth_loss_weight > 0:                           # MPI.py:617-623
            extra["a_smooth"] = (ts["smooth"][2] / nx + ts["smooth"][3] / ny).float().reshape(1, -1)
        if args.d_smooth_loss_weight > 0:                           # MPI.py:625-638
            disp = variables["disp_norm"]
            depth_grad = (disp[:, 1:, :-1] - disp[:, 1:, 1:]).abs() + (disp[:, :-1, 1:] - disp[:, 1:, 1:]).abs()
            rgb = rgbl[:, :3]
            edge = ((rgb[..., 1:, :-1] - rgb[..., 1:, 1:]).abs().sum(dim=1) + (rgb[..., :-1, 1:] - rgb[..., 1:, 1:]).abs().sum(dim=1))
            weight = (-edge * args.edge_scale + 1).clamp_min(0)
            extra["d_smooth"] = (depth_grad * weight).mean().reshape(1, -1)
        if getattr(args, "l_smooth_loss_weight", 0) > 0 and ts["l_smooth"] is not None:   # MPI.py:640-646
            extra["l_smooth"] = (ts["l_smooth"][0] / (3 * nx) + ts["l_smooth"][1] / (3 * ny)).float().reshape(1, -1)
        if args.density_loss_weight > 0:                            # MPI.py:648-651
            extra["density"] = (variables["alpha"] - 1).abs().mean().reshape(1, -1)
        return rgbl, extra

    # ------------------------------------------------------------------ optimiser (MPI.py:126-152)
    def get_optimizer(self):
        """Adam(betas=(0.9, 0.999)) over everything but `_verts` (its own group with lr * optimize_verts_gain); torch's
        default eps, unlike stage 2 (MPI.py:126-141)."""
        args = self.args
        named = dict(self.named_parameters())
        groups = [{'params': [v for k, v in named.items() if k != "_verts"]},
                  {'params': [named["_verts"]], 'lr': args.lrate * args.optimize_verts_gain}]
        if args.optimizer == 'adam':
            return torch.optim.Adam(params=groups, lr=args.lrate, betas=(0.9, 0.999))
        if args.optimizer == 'sgd':
            return torch.optim.SGD(params=groups, lr=args.lrate, momentum=0.9)
        raise RuntimeError(f"Unrecongnized optimizer type {args.optimizer}")

    def get_lrate(self, step):
        args = self.args
        scaling = 0.1 ** (step / (args.lrate_decay * 1000))
        return [("lr", args.lrate * scaling), ("vertlr", args.lrate * args.optimize_verts_gain * scaling)]

    def update_step(self, step):
        """MPI.py:154-157: geometry optimisation would start at `optimize_geo_start` (1e7 in every shipped config)."""
        if step >= getattr(self.args, "optimize_geo_start", 10000000):
            raise NotImplementedError("geometry optimisation (optimize_geo_start) is never reached by the shipped configs and is "
                                      "not supported: the kernels assume fronto-parallel planes of axis-aligned quads")

    def init_from_mpi(self, state_dict):
        """Load a stage-1 checkpoint written by `state_dict` — dense or culled (MPI.py:173-205)."""
        sd = state_dict
        dev = self.atlas.device
        self._verts.data = sd['_verts'].to(self._verts)
        self.uvs.data = sd['uvs'].to(self.uvs)
        self.atlas.data = ops.as_texels(sd['atlas'].to(device=dev, dtype=torch.float32))
        self.uvfaces.data = sd['uvfaces'].to(self.uvfaces)
        self.faces.data = sd['faces'].to(self.faces)
        self.ref_extrin.data = sd['ref_extrin'].to(self.ref_extrin)
        self.ref_intrin.data = sd['ref_intrin'].to(self.ref_intrin)
        self.planedepth.data = sd['planedepth'].to(self.planedepth)
        for k in self._SCALARS:
            setattr(self, k, sd["self." + k])
        if "atlas_mask" in sd and hasattr(self, "atlas_mask"):
            self.atlas_mask.data = sd["atlas_mask"].to(device=dev, dtype=torch.float32)
        if "self.has_dyn" in sd.keys():
            for k in self._SCALARS_DYN:
                setattr(self, k, sd["self." + k])
            self.register_parameter("uvs_dyn", nn.Parameter(sd['uvs_dyn'].to(self.uvs), requires_grad=True))
            self.register_buffer("uvfaces_dyn", sd['uvfaces_dyn'].to(self.uvfaces))
            self.register_buffer("faces_dyn", sd['faces_dyn'].to(self.faces))
            self.register_parameter("atlas_dyn", nn.Parameter(ops.as_texels(sd['atlas_dyn'].to(device=dev, dtype=torch.float32)),
                                                              requires_grad=True))
            if hasattr(self, "atlas_mask"):                         # a culled model has no loop mask any more (MPI.py:440-441)
                self.args.learn_loop_mask = False
                del self.atlas_mask
        self.invalidate_geometry()

    # ------------------------------------------------------------------ checkpoint format (MPI.py:207-221)
    _SCALARS = ("is_sparse", "atlas_full_w", "atlas_full_h", "atlas_grid_h", "atlas_grid_w")
    _SCALARS_DYN = ("has_dyn", "atlas_full_dyn_w", "atlas_full_dyn_h", "atlas_grid_dyn_h", "atlas_grid_dyn_w")

    def state_dict(self, destination=None, prefix='', keep_vars=False):
        """The tensors plus the layout scalars under "self.<name>" keys: what `MPMeshVid.init_from_mpi` (MPV.py:235-288)
        and the reference's own stage 2 load."""
        sd = super().state_dict()
        for k in self._SCALARS:
            sd["self." + k] = getattr(self, k)
        if hasattr(self, "atlas_dyn"):
            for k in self._SCALARS_DYN:
                sd["self." + k] = getattr(self, k)
        return sd

    # ------------------------------------------------------------------ tile culling (MPI.py:289-442)
    @staticmethod
    def _tile_grid(n, max_ratio=4):
        """Rows x columns of the packed atlas for n tiles and the number of filler tiles (MPI.py:367-381): among the row
        counts in [sqrt(n / max_ratio), sqrt(n)) the one that leaves the fewest empty cells when n // rows + 1 columns
        are used."""
        if n == 0:
            return 0, 0, 0
        rows = np.arange(int(np.sqrt(n / max_ratio)), int(np.sqrt(n)))
        if len(rows) == 0 or rows[0] == 0:
            raise ValueError(f"too few tiles ({n}) to pack an atlas (the reference's packing needs n >= 16)")
        h = int(rows[np.argmin(rows - n % rows)])
        w = n // h + 1
        return h, w, h * w - n

    @staticmethod
    def _morph(x, n_erode, n_dilate):
        """3x3 erosion (zero padding: the border erodes) n_erode times, then 3x3 dilation n_dilate times (utils.py:298-317)."""
        for _ in range(n_erode):
            x = -F.max_pool2d(-F.pad(x, (1, 1, 1, 1), value=0.0), 3, stride=1)
        for _ in range(n_dilate):
            x = F.max_pool2d(x, 3, stride=1, padding=1)
        return x

    @torch.no_grad()
    def sparsify_faces(self, erode_num=2, alpha_thresh=0.03, loop_thresh=0.5):
        """Tile culling (MPI.py:289-442): every quad's atlas tile is resampled to a private (th x tw)-texel tile; quads whose
        (eroded / dilated) alpha never exceeds `alpha_thresh` are dropped; the kept ones whose (eroded / dilated) loop mask
        exceeds `loop_thresh` become DYNAMIC tiles, the rest static; both sets are packed row-major into new atlases (last
        tile repeated as filler) with four private uv corners per tile.  Afterwards `has_dyn` / `is_sparse` are set,
        `atlas_mask` is gone and `learn_loop_mask` is off — the state `MPMeshVid.init_from_mpi` expects."""
        if self.is_sparse:
            raise Vl3dError("sparsify_faces: the model has been culled already")
        if not hasattr(self, "atlas_mask"):
            raise Vl3dError("sparsify_faces needs the loop-mask atlas (learn_loop_mask)")
        if getattr(self.args, "sparsify_rmfirstlayer", 0) > 0:
            raise NotImplementedError("sparsify_rmfirstlayer > 0 is used by no shipped config and is not supported")
        atlas = self.atlas.data.contiguous()                                   # standard NCHW for the resampling below
        Ha, Wa = atlas.shape[-2:]
        corners = self.uvfaces.reshape(-1, 6)                                  # (v00, v01, v11, v11, v10, v00) per quad
        assert bool((corners[:, 0] == corners[:, 5]).all()) and bool((corners[:, 2] == corners[:, 3]).all())
        uvs = self.uvs.data
        ext_w = (uvs[corners[0, 1]] - uvs[corners[0, 0]])[0].item()            # uv extent of a quad's tile
        ext_h = (uvs[corners[0, 4]] - uvs[corners[0, 0]])[1].item()
        tw, th = int(np.round(ext_w / 2 * (Wa - 1))), int(np.round(ext_h / 2 * (Ha - 1)))
        oy, ox = torch.meshgrid(torch.linspace(0, ext_h, th), torch.linspace(0, ext_w, tw), indexing="ij")
        offs = torch.stack([ox, oy], dim=-1)[None].to(atlas)                    # (1, th, tw, 2)
        nq = len(corners)
        grid = (uvs[corners[:, 0]][:, None, None, :] + offs).reshape(1, nq * th, tw, 2)
        # untouched texels (still at the initial logit) count as empty
        a_logit = atlas[:, 3:4].clone()
        a_logit[a_logit == ALPHA_INIT_VAL] = -10
        l_logit = self.atlas_mask.data.clone()
        l_logit[l_logit == ALPHA_INIT_VAL] = -10
        alpha = self._morph(torch.sigmoid(a_logit), erode_num, erode_num + 2)
        loop = self._morph(torch.sigmoid(l_logit), erode_num, erode_num)
        sample = lambda img: F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True) \
            .reshape(img.shape[1], nq, th, tw).permute(1, 2, 3, 0)              # (nq, th, tw, C)
        tiles_rgba = sample(atlas)
        keep = sample(alpha).reshape(nq, -1).max(dim=-1)[0] > alpha_thresh
        dyn = keep & (sample(loop).reshape(nq, -1).max(dim=-1)[0] > loop_thresh)
        sta = keep & ~dyn
        quad_faces = self.faces.reshape(-1, 2, 3)

        def pack(mask):
            n = int(mask.sum())
            if n == 0:
                raise ValueError("sparsify_faces: no static or no dynamic tile survives (the reference cannot pack that either)")
            h, w, filler = self._tile_grid(n)
            t = tiles_rgba[mask]
            t = torch.cat([t, t[-1:].expand(filler, -1, -1, -1)])
            new_atlas = t.reshape(h, w, th, tw, -1).permute(4, 0, 2, 1, 3).reshape(1, -1, h * th, w * tw)
            ah, aw = new_atlas.shape[-2:]
            # four private corners per tile; the column step uses `th` like the reference (MPI.py:410): tiles are square
            sy, sx = 2 / (ah - 1) * (th - 1), 2 / (aw - 1) * (tw - 1)
            corner = torch.tensor([[0, 0], [sx, 0], [0, sy], [sx, sy]]).to(uvs)
            v0, u0 = torch.meshgrid(torch.arange(0, ah, th) / (ah - 1) * 2 - 1, torch.arange(0, aw, th) / (aw - 1) * 2 - 1,
                                    indexing="ij")
            uv0 = torch.stack([u0, v0], dim=-1).to(uvs)
            quv = (uv0[:, :, None, :] + corner[None, None]).reshape(-1, 4, 2)[:n]
            uvf = (torch.arange(n)[:, None, None] * 4 + torch.tensor([[0, 1, 3], [3, 2, 0]])[None]).to(self.uvfaces)
            return new_atlas, quv.reshape(-1, 2), uvf.reshape(-1, 3).long(), quad_faces[mask].reshape(-1, 3).long(), (h, w)

        atlas_s, uvs_s, uvf_s, faces_s, grid_s = pack(sta)
        atlas_d, uvs_d, uvf_d, faces_d, grid_d = pack(dyn)
        self.is_sparse = True
        self.atlas_grid_h, self.atlas_grid_w = grid_s
        self.atlas_full_h, self.atlas_full_w = atlas_s.shape[-2:]
        self.atlas_grid_dyn_h, self.atlas_grid_dyn_w = grid_d
        self.atlas_full_dyn_h, self.atlas_full_dyn_w = atlas_d.shape[-2:]
        self.register_parameter("uvs", nn.Parameter(uvs_s, requires_grad=True))
        self.register_buffer("uvfaces", uvf_s)
        self.register_buffer("faces", faces_s)
        self.register_parameter("atlas", nn.Parameter(ops.as_texels(atlas_s), requires_grad=True))
        self.has_dyn = True
        self.register_parameter("uvs_dyn", nn.Parameter(uvs_d, requires_grad=True))
        self.register_buffer("uvfaces_dyn", uvf_d)
        self.register_buffer("faces_dyn", faces_d)
        self.register_parameter("atlas_dyn", nn.Parameter(ops.as_texels(atlas_d), requires_grad=True))
        self.args.learn_loop_mask = False
        del self.atlas_mask
        self.invalidate_geometry()
        return dict(quads=nq, kept=int(keep.sum()), dynamic=int(dyn.sum()), tile=(th, tw))
