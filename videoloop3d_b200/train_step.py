"""The stage-2 optimisation step (reference: train_3dvid.py:214-255 `run_iter`).

Two front ends over the same kernels:

* `make_run_iter(args, nerf, device)` — reference-shaped closure `run_iter(stepi, optimizer_, datainfo_)`
  that goes through `nerf(...)`, autograd and `optimizer_.step()` exactly like the reference loop, so
  the surrounding `train()` code can stay as it is.
* `FusedLoopStep` — the same step without autograd bookkeeping: persistent buffers, no host
  synchronisation, one stream; optionally T-sharded over `torch.distributed` ranks
  (SURVEY.md §8(e)): every rank renders / back-propagates / optimises its own contiguous block of
  frames, rendered frames are all-gathered (NCCL) for the temporal patches, static-tile gradients are
  all-reduced.  This is what bench.py times.
"""
from __future__ import annotations

from argparse import Namespace

import numpy as np
import torch

import ctypes as C
import os

from . import _lib, ops, schedule, tiles
from ._lib import Vl3dError
from .loop_loss import Patch3DGPNNDirectLoss, Patch3DGPNNLowMemLoss, _check_dist
from .mpv import MPMeshVid, pose2extrin_torch


def default_args(**overrides):
    """Namespace with the flags the hot path reads; defaults = config_parser.py defaults overlaid with
    configs/mpv_base.txt (the config every shipped stage-2 scene uses)."""
    a = dict(
        mpv_frm_num=50, mpv_isloop=True, init_std=0.02, mpi_h_scale=1.1, mpi_w_scale=1.1, mpi_h_verts=27,
        mpi_w_verts=48, mpi_d=32, atlas_grid_h=4, atlas_size_scale=1, atlas_cnl=4, normalize_verts=False, fp16=False,
        rgb_mlp_type="direct", rgb_activate="sigmoid", alpha_activate="sigmoid", bg_color="",
        scale_invariant=True, add_intrin_noise=True, add_uv_noise=False,
        sparsity_loss_weight=0.0, rgb_smooth_loss_weight=0.2, a_smooth_loss_weight=0.2, density_loss_weight=0.0,
        d_smooth_loss_weight=0.0, l_smooth_loss_weight=0.0,
        optimizer="adam", lrate=0.5, lrate_decay=100, lrate_adaptive=True, optimize_verts_gain=1.0,
        optimize_geo_start=10000000,
        swd_macro_block=65, swd_loss_gain_ref=3.5, loss_name_ref="gpnn_lm", swd_alpha_ref=0.0, swd_patch_size_ref=11,
        swd_patcht_size_ref=3, swd_stride_ref=4, swd_stridet_ref=1, swd_dist_fn_ref="mse", swd_rou_ref="-2",
        swd_scaling_ref=0.1, swd_factor_ref=1,
        loss_name="gpnn_lm", swd_alpha=10000.0, swd_patch_size=3, swd_patcht_size=3, swd_stride=2, swd_stridet=1,
        swd_dist_fn="mse", swd_rou="-2", swd_scaling=0.1, swd_factor=1,
        patch_h_size=180, patch_w_size=320, patch_h_stride=90, patch_w_stride=160, i_img=20, i_print=10, gpu_num=1,
    )
    a.update(overrides)
    return Namespace(**a)


def default_args_stage1(**overrides):
    """Namespace with the flags `MPMesh` reads; defaults = config_parser.py defaults overlaid with configs/mpi_base.txt."""
    a = dict(
        mpi_h_scale=1.6, mpi_w_scale=1.6, mpi_h_verts=36, mpi_w_verts=64, mpi_d=32, atlas_grid_h=4, atlas_size_scale=1,
        normalize_verts=False, upsample_stage="", learn_loop_mask=True, rgb_mlp_type="direct", rgb_activate="sigmoid",
        alpha_activate="sigmoid", bg_color="", add_uv_noise=False, normalize_blendweight_fordepth=False, edge_scale=4.0,
        sparsity_loss_weight=0.004, rgb_smooth_loss_weight=0.2, a_smooth_loss_weight=0.5, density_loss_weight=0.02,
        d_smooth_loss_weight=0.0, l_smooth_loss_weight=0.0, scale_invariant=True, add_intrin_noise=True,
        optimizer="adam", lrate=0.05, lrate_decay=100, optimize_verts_gain=1.0,
        patch_h_size=180, patch_w_size=320, patch_h_stride=90, patch_w_stride=160,
    )
    a.update(overrides)
    return Namespace(**a)


def loss_config(args, ref_view: bool):
    """The two per-view loss configs built in train_3dvid.py:163-190."""
    if ref_view:
        return dict(loss_name=args.loss_name_ref, loss_gain=args.swd_loss_gain_ref, patch_size=args.swd_patch_size_ref,
                    patcht_size=args.swd_patcht_size_ref, stride=args.swd_stride_ref, stridet=args.swd_stridet_ref,
                    alpha=args.swd_alpha_ref, rou=args.swd_rou_ref, scaling=args.swd_scaling_ref,
                    dist_fn=args.swd_dist_fn_ref, macro_block=args.swd_macro_block, factor=args.swd_factor_ref)
    return dict(loss_name=args.loss_name, patch_size=args.swd_patch_size, patcht_size=args.swd_patcht_size,
                stride=args.swd_stride, stridet=args.swd_stridet, alpha=args.swd_alpha, rou=args.swd_rou,
                scaling=args.swd_scaling, dist_fn=args.swd_dist_fn, macro_block=args.swd_macro_block,
                factor=args.swd_factor)


def make_run_iter(args, nerf, device, writer=None, on_log=None):
    """Reference-shaped `run_iter(stepi, optimizer_, datainfo_)` (train_3dvid.py:214-255).
    `nerf` is the (DataParallel-like) wrapper or the bare module."""

    def run_iter(stepi, optimizer_, datainfo_):
        datainfo_ = [d.to(device) if torch.is_tensor(d) else d for d in datainfo_]
        h_starts, w_starts, b_pose, b_intrin, b_rgbs, loss_cfg = datainfo_
        if b_rgbs.dtype == torch.uint8:                             # MVVidPatchDataset(storage="uint8"): bytes until here
            b_rgbs = ops.u8_to_unit(b_rgbs[0])[None]
        b_extrin = pose2extrin_torch(b_pose)
        patch_h, patch_w = b_rgbs.shape[-2:]
        if args.add_intrin_noise:
            dxy = torch.rand(2).type_as(b_intrin) - 0.5          # half pixel (train_3dvid.py:222-225)
            b_intrin = b_intrin.clone()
            b_intrin[:, :2, 2] += dxy
        nerf.train()
        rgb, extra = nerf(patch_h, patch_w, b_extrin, b_intrin, res=b_rgbs, losscfg=loss_cfg)
        swd_loss = extra.pop("swd").mean()
        args_var = vars(args)
        extra_losses = {k: v.mean() * args_var[f"{k}_loss_weight"] for k, v in extra.items()
                        if args_var[f"{k}_loss_weight"] > 0}
        loss = swd_loss
        for v in extra_losses.values():
            loss = loss + v
        optimizer_.zero_grad()
        loss.backward()
        optimizer_.step()
        if on_log is not None:
            on_log(stepi, loss, swd_loss, extra_losses)
        return loss.detach()

    return run_iter


def make_run_iter_stage1(args, nerf, device, on_log=None):
    """Reference-shaped stage-1 `run_iter(stepi, optimizer_, datainfo_)` (train_3d.py:189-236) around an `MPMesh`:
    `datainfo_ = (h_start, w_start, pose (1,3,4), intrin (1,3,3), rgb (1,3,h,w), loopmask (1,h,w))`; loss = scale-invariant
    MSE + cross-entropy of the rendered loop-mask label + the weighted extra terms.  The render and every regulariser run
    in libvl3d; the image / label losses are a handful of element-wise ops on one (1,4,h,w) image."""

    def run_iter(stepi, optimizer_, datainfo_):
        datainfo_ = [d.to(device) if torch.is_tensor(d) else d for d in datainfo_]
        h_starts, w_starts, b_pose, b_intrin, b_rgbs, b_loopmask = datainfo_
        b_extrin = pose2extrin_torch(b_pose)
        patch_h, patch_w = b_rgbs.shape[-2:]
        if args.add_intrin_noise:
            dxy = torch.rand(2).type_as(b_intrin) - 0.5          # half pixel (train_3d.py:194-197)
            b_intrin = b_intrin.clone()
            b_intrin[:, :2, 2] += dxy
        nerf.train()
        rgb, extra = nerf(patch_h, patch_w, b_extrin, b_intrin)
        loop_loss = 0
        if args.learn_loop_mask:                                  # train_3d.py:201-212
            label = torch.clamp(rgb[:, -1], 0.001, 1 - 0.001)
            loop_loss = -(b_loopmask * torch.log(label) + (1 - b_loopmask) * torch.log(1 - label)).mean()
            rgb = rgb[:, :3]
        if args.scale_invariant:                                  # train_3d.py:217-220
            scale = torch.exp(torch.log((b_rgbs + 0.01) / (rgb.detach() + 0.01)).mean())
            rgb = rgb * ((scale + 3) / 4)
        img_loss = torch.mean((rgb - b_rgbs) ** 2)
        args_var = vars(args)
        extra_losses = {k: v.mean() * args_var[f"{k}_loss_weight"] for k, v in extra.items()
                        if args_var[f"{k}_loss_weight"] > 0}
        loss = img_loss + loop_loss
        for v in extra_losses.values():
            loss = loss + v
        optimizer_.zero_grad()
        loss.backward()
        optimizer_.step()
        if on_log is not None:
            on_log(stepi, loss, img_loss, extra_losses)
        return loss.detach()

    return run_iter


def partition(n, world):
    """Contiguous blocks [b[r], b[r+1]) of n units over `world` ranks."""
    return [(n * r) // world for r in range(world + 1)]


def owned_frame_ranges(bounds, rank, T, pad):
    """Frame ranges of the padded video x = cat(rgb, rgb[:pad]) whose loss gradient a rank needs:
    its own frames [t0,t1) and the looped copies T+t of the frames t < pad it owns (MPV.py:490-492)."""
    t0, t1 = bounds[rank], bounds[rank + 1]
    out = [(t0, t1)]
    if t0 < pad:
        out.append((T + t0, T + min(t1, pad)))
    return out


def gather_frames(video, bounds, rank, T, group, async_op=False):
    """All-gather of the frame blocks of `video[:T]` (rank r holds video[bounds[r]:bounds[r+1]] and receives the
    rest in place).  Equal blocks: one in-place `all_gather_into_tensor`; ragged blocks: list all-gather.
    Returns the work handle when `async_op`."""
    import torch.distributed as dist
    t0, t1 = bounds[rank], bounds[rank + 1]
    if len({b - a for a, b in zip(bounds[:-1], bounds[1:])}) == 1:
        return dist.all_gather_into_tensor(video[:T], video[t0:t1], group=group, async_op=async_op)
    parts = [video[a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    return dist.all_gather(parts, video[t0:t1].clone(), group=group, async_op=async_op)


def exchange_row_bands(nn, rows, rank, group):
    """Every rank filled the rows [rows[rank], rows[rank+1]) of the NN index map `nn` (ho, wo, n1); afterwards
    all ranks hold the whole map.  Equal bands: in-place all-gather; ragged bands: the caller zeroed the rest
    and the disjoint bands are summed."""
    import torch.distributed as dist
    if rows_equal(rows):
        dist.all_gather_into_tensor(nn, nn[rows[rank]:rows[rank + 1]], group=group)
    else:
        dist.all_reduce(nn, group=group)


def rows_equal(rows):
    return len({b - a for a, b in zip(rows[:-1], rows[1:])}) == 1


def band_layout(h, p, s, ho, world):
    """Row bands of the looping loss for `world` ranks (patch positions are independent, utils_vid.py:211-215).

    Rank r searches the patch rows [pr0, pr1) and owns the loss / gradient of the pixel rows [own0, own1) (a partition
    of [0, h)).  A pixel row is covered by up to ceil(p/s) patch rows, so the vote of the first owned rows also needs
    the matches of the `halo` patch rows above pr0: the rank's band buffer holds the pixel rows [ya, yb) = everything
    its patch rows [pr0 - halo, pr1) and its owned rows touch.  Returns a list of dicts (one per rank)."""
    rows = partition(ho, world)
    M = (p - 1) // s
    out = []
    for r in range(world):
        pr0, pr1 = rows[r], rows[r + 1]
        halo = min(M, pr0)
        own0 = pr0 * s if r > 0 else 0
        own1 = rows[r + 1] * s if r < world - 1 else h
        ya = (pr0 - halo) * s
        yb = max((pr1 - 1) * s + p if pr1 > pr0 else ya, own1)
        out.append(dict(pr0=pr0, pr1=pr1, halo=halo, own0=own0, own1=own1, ya=ya, yb=min(yb, h)))
    return out


def frames_to_bands(frames_local, band_out, bounds, bands, rank, group, send_ws=None):
    """All-to-all #1: every rank holds ALL rows of ITS frames (`frames_local`: (Tl,C,h,w)); afterwards it holds ITS
    row band of ALL frames (`band_out[:T]`: (>=T,C,hb,w), frame-major, so the block received from rank q lands
    contiguously at frames [bounds[q], bounds[q+1])).  Returns the packed send buffer (reusable workspace)."""
    import torch.distributed as dist
    world = len(bands)
    Tl, C_, _, w = frames_local.shape
    T = bounds[-1]
    hb = bands[rank]["yb"] - bands[rank]["ya"]
    in_splits = [Tl * C_ * (b["yb"] - b["ya"]) * w for b in bands]
    out_splits = [(bounds[q + 1] - bounds[q]) * C_ * hb * w for q in range(world)]
    n_send = sum(in_splits)
    if send_ws is None or send_ws.numel() < n_send or send_ws.dtype != frames_local.dtype or send_ws.device != frames_local.device:
        send_ws = torch.empty(n_send, dtype=frames_local.dtype, device=frames_local.device)
    off = 0
    for q, b in enumerate(bands):                                   # pack: rank q's rows of my frames, contiguous
        n = in_splits[q]
        send_ws[off:off + n].view(Tl, C_, b["yb"] - b["ya"], w).copy_(frames_local[:, :, b["ya"]:b["yb"]])
        off += n
    dist.all_to_all_single(band_out[:T].reshape(-1), send_ws[:n_send], out_splits, in_splits, group=group)
    return send_ws


def bands_to_frames(band_grad, frames_grad_out, bounds, bands, rank, group, send_ws=None, recv_ws=None):
    """All-to-all #2 (the adjoint of #1 restricted to owned rows): `band_grad` (>=T,C,hb,w) holds dL/dx for my band of
    all frames; every rank receives, for ITS frames, the rows each rank owns -> `frames_grad_out` (Tl,C,h,w)."""
    import torch.distributed as dist
    world = len(bands)
    me = bands[rank]
    Tl, C_, h, w = frames_grad_out.shape
    lo, hi = me["own0"] - me["ya"], me["own1"] - me["ya"]
    in_splits = [(bounds[q + 1] - bounds[q]) * C_ * (hi - lo) * w for q in range(world)]
    out_splits = [Tl * C_ * (b["own1"] - b["own0"]) * w for b in bands]
    n_send, n_recv = sum(in_splits), sum(out_splits)
    kw = dict(dtype=band_grad.dtype, device=band_grad.device)
    if send_ws is None or send_ws.numel() < n_send:
        send_ws = torch.empty(n_send, **kw)
    if recv_ws is None or recv_ws.numel() < n_recv:
        recv_ws = torch.empty(n_recv, **kw)
    T = bounds[-1]
    send_ws[:n_send].view(T, C_, hi - lo, w).copy_(band_grad[:T, :, lo:hi])     # frames are already in rank order
    dist.all_to_all_single(recv_ws[:n_recv], send_ws[:n_send], out_splits, in_splits, group=group)
    off = 0
    for q, b in enumerate(bands):
        n = out_splits[q]
        frames_grad_out[:, :, b["own0"]:b["own1"]].copy_(recv_ws[off:off + n].view(Tl, C_, b["own1"] - b["own0"], w))
        off += n
    return send_ws, recv_ws


class PeerExchange:
    """The band-sharded loss's exchanges over NVLink peer memory instead of NCCL all-to-alls.

    Every rank owns one symmetric-memory allocation (torch.distributed._symmetric_memory: the same virtual layout on every
    rank, each rank's copy mapped into all the others) that holds its row band of the rendered video, the NN index map
    and the loss gradient of its own frames.  A rank then STORES its rows of its frames directly into the band buffers of
    the ranks that need them (`vl3d_copy_boxes`: one launch, no staging copies, no collective call), and later its band's
    gradient rows into the frame owners' gradient buffers; a device-side barrier on the symmetric allocation's signal
    pads (stream-ordered, no host sync) separates writers from readers.  Buffer reuse across steps is safe with the three
    barriers of a step: a rank can only write into a peer's band (NN map, gradient) buffer of step k+1 after that peer has
    passed the barrier that follows its last read of step k's content."""

    def __init__(self, group, device):
        import torch.distributed as dist
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.key, self.buf, self.hdl = None, None, None

    def layout(self, T, pad, h, w, bands, bounds, nn_shape):
        """(Re)allocate for this problem shape (collective: shapes are the same on every rank).  Word offsets of the three
        regions inside the allocation; every region is sized for the largest rank so that the layout is symmetric."""
        key = (T, pad, h, w, tuple((b["ya"], b["yb"]) for b in bands), tuple(bounds), tuple(nn_shape))
        if key != self.key:
            import torch.distributed._symmetric_memory as symm
            hb_max = max(b["yb"] - b["ya"] for b in bands)
            tl_max = max(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:]))
            n_x = (T + pad) * 3 * hb_max * w
            n_nn = int(np.prod(nn_shape))
            n_g = tl_max * 3 * h * w
            al = lambda n: (n + 63) // 64 * 64
            self.off = (0, al(n_x), al(n_x) + al(n_nn))
            self.buf = symm.empty(al(n_x) + al(n_nn) + al(n_g), dtype=torch.float32, device=self.device)
            self.hdl = symm.rendezvous(self.buf, self.group)
            self.key = key
            self.bands, self.bounds, self.dims, self.nn_shape = bands, bounds, (T, pad, h, w), tuple(nn_shape)
        return self

    def _region(self, rank, which, shape):
        n = int(np.prod(shape))
        if rank == self.rank:
            return self.buf[self.off[which]:self.off[which] + n].view(shape)
        return self.hdl.get_buffer(rank, (n,), torch.float32, self.off[which]).view(shape)

    def x_band(self, rank):
        T, pad, h, w = self.dims
        b = self.bands[rank]
        return self._region(rank, 0, (T + pad, 3, b["yb"] - b["ya"], w))

    def nn(self, rank):
        return self._region(rank, 1, self.nn_shape).view(torch.int32)

    def grad_frames(self, rank):
        T, pad, h, w = self.dims
        return self._region(rank, 2, (self.bounds[rank + 1] - self.bounds[rank], 3, h, w))

    def barrier(self):
        self.hdl.barrier(channel=0)

    def frames_to_bands(self, frames_local):
        """My frames (Tl,3,h,w) -> every rank's band buffer (frames [t0,t1) and, for t < pad, their looped copies T + t)."""
        T, pad, h, w = self.dims
        t0, t1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        boxes = []
        for q, b in enumerate(self.bands):
            dst = self.x_band(q)
            src = frames_local[:, :, b["ya"]:b["yb"]]
            boxes.append((src, dst[t0:t1], None))
            if t0 < pad:
                n = min(t1, pad) - t0
                boxes.append((src[:n], dst[T + t0:T + t0 + n], None))
        ops.copy_boxes(boxes)
        self.barrier()

    def share_nn(self, rows, everything=True):
        """My patch rows [rows[rank], rows[rank+1]) of the NN map -> the other ranks' maps: all of them (`everything`: every
        rank ends up with the whole map), or only the halo rows a rank's vote needs from the ranks above it."""
        r0, r1 = rows[self.rank], rows[self.rank + 1]
        mine = self.nn(self.rank)
        boxes = []
        for q, b in enumerate(self.bands):
            if q == self.rank:
                continue
            a0, a1 = (r0, r1) if everything else (max(r0, b["pr0"] - b["halo"]), min(r1, b["pr0"]))
            if a1 > a0:
                boxes.append((mine[a0:a1].view(torch.float32)[None], self.nn(q)[a0:a1].view(torch.float32)[None], None))
        ops.copy_boxes(boxes)
        self.barrier()

    def bands_to_frames(self, g_band):
        """dL/dx of my band's owned rows, all frames -> the frame owners' gradient buffers (loop-pad adjoint folded in)."""
        T, pad, h, w = self.dims
        me = self.bands[self.rank]
        lo, hi = me["own0"] - me["ya"], me["own1"] - me["ya"]
        boxes = []
        for q in range(self.world):
            b0, b1 = self.bounds[q], self.bounds[q + 1]
            dst = self.grad_frames(q)[:, :, me["own0"]:me["own1"]]
            nf = max(0, min(b1, pad) - b0)                           # frames of q that have a looped copy
            if nf > 0:
                boxes.append((g_band[b0:b0 + nf, :, lo:hi], dst[:nf], g_band[T + b0:T + b0 + nf, :, lo:hi]))
            if b1 - b0 > nf:
                boxes.append((g_band[b0 + nf:b1, :, lo:hi], dst[nf:], None))
        ops.copy_boxes(boxes)
        self.barrier()


class FusedLoopStep:
    """render + looping loss + backward + Adam for one (view, patch) item, fused and sync-free.

    step(h, w, tar_extrin (1,4,4), tar_intrin (1,3,3), res (1,F,3,h,w) on device, losscfg (un-batched dict), lr)
    returns a dict of device scalars {'loss','swd','rgb_smooth','a_smooth'} (no host sync).

    With `group` (a torch.distributed process group, world size G) the work is sharded:
      * frames: rank r owns the contiguous block [t0,t1) of `atlas_dyn` — it renders, back-propagates and
        optimises only those frames (parameters / gradients / Adam state of other frames never move);
        `global_frames=T` means the model on this rank holds ONLY its own block (memory-sharded);
      * one all-gather of the rendered frames (temporal patches, loop pad and the scale-invariant mean need
        the whole video), patch rows of the NN search split across ranks + one all-reduce of the int32 index
        map, the loss/regulariser partial sums all-reduced as 5 doubles;
      * static-tile gradients all-reduced (SUM) — the single gradient collective.
    """

    def __init__(self, model: MPMeshVid, group=None, betas=(0.9, 0.999), eps=6e-8, global_frames=None, timers=False,
                 overlap_chunks=1, fused=None, fused_opts=None, loss_shard="rows", exchange=None, gather_nn=True):
        """`fused`: how backward + Adam of the dynamic atlas run —
        "off": separate kernels (zero-fill, vl3d_composite_bwd, vl3d_adam_step);
        "generic": one persistent kernel, tiles of chunk c interleaved with Adam of chunk c-1 (any layout);
        "band" / "band-zero": the same kernel walking the screen in row bands so that a band's gradient rows stay
        L2-resident between accumulation and Adam (dense layout only; "-zero": rows are zeroed just ahead of the
        band and dropped after Adam instead of living in HBM as zeros);
        "own": the generic queue in owner mode (`vl3d_fused_bwd_adam_own`, dense layout + regulariser): texels met by the
        pixels of one screen tile only are optimised inside that tile and never cross HBM as gradients — correct and
        tested, but latency-bound and slower on B200 (profiles/r02_fused_bwd_adam.md);
        None: VL3D_FUSED from the environment (read once, here), else "auto" = generic, except band for a dense model when
        this rank owns at most 6 frame chunks of a large image.  (On B200 the band schedules do NOT keep the gradient on
        chip under the DRAM-saturating Adam traffic — profiles/r02_fused_bwd_adam.md — but with Adam lagging one wave
        of resident tiles behind they match the generic schedule in steady state at 720p (42.9 vs 43.7 ms) and shorten
        its pipeline fill / drain from one frame chunk to ~1/8 of one, which matters when a rank owns few chunks; at
        180x320 / 360x640 the finer dependencies cost more than they save: 4.3 vs 3.2 ms / 12.0 vs 11.4 ms.)"""
        if not model.atlas_dyn.is_cuda:
            raise Vl3dError("FusedLoopStep needs the model on a CUDA device")
        self.model = model
        self.group = group
        self.betas, self.eps = betas, eps
        self.t = 0                       # Adam step counter (fresh optimiser per pyramid level: call reset())
        self._buf = {}
        self._state = {}
        self.world, self.rank = 1, 0
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.local_model = global_frames is not None
        T = int(global_frames) if self.local_model else model.atlas_dyn.shape[0]
        self.T = T
        self.bounds = partition(T, self.world)
        self.t0, self.t1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        if self.t1 <= self.t0:
            raise Vl3dError(f"rank {self.rank} owns no frames (T={T}, world={self.world})")
        if self.local_model and model.atlas_dyn.shape[0] != self.t1 - self.t0:
            raise Vl3dError(f"memory-sharded model must hold {self.t1 - self.t0} frames, has {model.atlas_dyn.shape[0]}")
        self.timers = {} if timers else None
        # optional: pipeline backward / Adam over `overlap_chunks` frame chunks on two streams.  Measured on
        # B200 at 720p: no gain (both contend for the L2/HBM path; the backward slows down by what Adam
        # gains), so the default is 1 (off).
        self.overlap_chunks = max(1, int(overlap_chunks))
        self._side = torch.cuda.Stream(device=model.atlas_dyn.device)
        # with several ranks: "rows" = the looping loss is sharded by pixel-row bands (two all-to-alls: rendered frames ->
        # row bands, loss gradient -> frames; every rank needs only its band of the target video); "frames" = the first
        # design (all-gather of the rendered frames, every rank holds the whole target video)
        self.loss_shard = os.environ.get("VL3D_LOSS_SHARD", loss_shard)
        if self.loss_shard not in ("rows", "frames"):
            raise ValueError(f"loss_shard={self.loss_shard!r}")
        # exchanges of the band-sharded loss: "p2p" = stores into the peers' symmetric-memory buffers (PeerExchange),
        # "nccl" = all-to-all collectives; None: VL3D_EXCHANGE, else p2p when symmetric memory can be set up.
        # gather_nn: give every rank the whole NN index map (`_buf["nn"]`); False (p2p only): a rank receives just the halo
        # patch rows its vote needs from the ranks above it.
        self.exchange = (exchange or os.environ.get("VL3D_EXCHANGE", "auto")).lower()
        if self.exchange not in ("p2p", "nccl", "auto"):
            raise ValueError(f"exchange={self.exchange!r}")
        self.gather_nn = bool(gather_nn)
        # The gradient buffer of the fused kernel lives in COMPRESSIBLE device memory when the device offers it (placement
        # only; VL3D_GRAD_COMPRESS=0 disables): two of the gradient's four HBM crossings per step are all-zero lines
        # (written back by the Adam items, fetched again by the first RED), which then move compressed: 253 -> 218 GB and
        # 43.2 -> 41.2 ms for the fused kernel at 720p.
        self.grad_compress = os.environ.get("VL3D_GRAD_COMPRESS", "1") != "0"
        self.grad_compressed = False
        self._peer = None
        self.fused = (fused or os.environ.get("VL3D_FUSED", "auto")).lower()
        if self.fused not in ("off", "generic", "band", "band-zero", "own", "auto"):
            raise ValueError(f"fused={self.fused!r}")
        self.fused_opts = dict(fused_opts or {})
        for k, cast in (("ctas_per_sm", int), ("row_block", int), ("zero_ahead", int), ("adam_lag", int), ("seg_texels", int)):
            e = os.environ.get("VL3D_FUSED_" + k.upper())
            if e is not None and k not in self.fused_opts:
                self.fused_opts[k] = cast(e)
        self._sched_cache = {}
        self.last_schedule = None

    def _timed(self, name):
        return _Timed(self.timers, name)

    def timer_ms(self, skip=0):
        """{kernel name: mean ms} from the CUDA events recorded so far (call after a synchronize)."""
        out = {}
        for k, evs in (self.timers or {}).items():
            evs = evs[skip:]
            if evs:
                out[k] = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
        return out

    def reset(self):
        """Fresh optimiser state (the reference builds a new Adam per pyramid level, train_3dvid.py:265)."""
        self.t = 0
        self._state = {}

    def _get(self, key, shape, dtype):
        b = self._buf.get(key)
        if b is None or tuple(b.shape) != tuple(shape) or b.dtype != dtype:
            b = torch.empty(shape, dtype=dtype, device=self.model.atlas_dyn.device)
            self._buf[key] = b
        return b

    def _like(self, key, ref):
        b = self._buf.get(key)
        if b is None or b.shape != ref.shape or tuple(b.stride()) != tuple(ref.stride()):
            b = torch.empty_like(ref)
            self._buf[key] = b
        return b

    def _adam_lag(self, w, smooth):
        """Tile rows the queue advances while one wave of resident tiles runs (+1): an Adam item placed that far behind
        the last tile row it depends on is claimed when those tiles are done, so no CTA spins on a counter."""
        gx = schedule.tile_grid(1, w, smooth)[0]
        sms = torch.cuda.get_device_properties(self.model.atlas_dyn.device).multi_processor_count
        per_sm = self.fused_opts.get("ctas_per_sm", 0) or 3
        return -(-sms * per_sm // gx) + 1

    def _schedule_for(self, mode, view, pack, h, w, smooth, dyn_hw):
        """Work-item table of the fused kernel for this view (cached: a table depends on the view only)."""
        key = (mode, bytes(view), bool(smooth), tuple(dyn_hw))
        sched = self._sched_cache.get(key)
        if sched is None:
            o = self.fused_opts
            if mode == "generic":
                sched = schedule.generic_schedule(h, w, dyn_hw[0], dyn_hw[1], smooth, seg_texels=o.get("seg_texels", 32768))
            elif mode == "own":
                homs = np.ctypeslib.as_array(view.hom).reshape(-1)[:pack.D * 9].reshape(pack.D, 9)
                hinv, reach, rect, cells = tiles.own_descriptor(pack.table, pack.D, pack.qh, pack.qw, homs, view.cx, view.cy, h, w,
                                                                dyn_hw, max_reach=o.get("max_reach", 3.0))
                sched = schedule.generic_schedule(h, w, dyn_hw[0], dyn_hw[1], smooth, seg_texels=o.get("seg_texels", 32768),
                                                  cells=cells)
                own = _lib.Own()
                C.memmove(own.hinv, np.ascontiguousarray(hinv).ctypes.data, hinv.nbytes)
                C.memmove(own.reach, np.ascontiguousarray(reach).ctypes.data, reach.nbytes)
                C.memmove(own.rect, np.ascontiguousarray(rect).ctypes.data, rect.nbytes)
                sched.own = own
            else:
                homs = np.ctypeslib.as_array(view.hom).reshape(-1)[:pack.D * 9].astype(np.float64)
                sched = schedule.band_schedule(homs, view.cx, view.cy, h, w, pack.table, pack.D, pack.qh, pack.qw, dyn_hw[0],
                                               dyn_hw[1], smooth, row_block=o.get("row_block", 8),
                                               zero_ahead=o.get("zero_ahead", 2), adam_lag=o.get("adam_lag", self._adam_lag(w, smooth)),
                                               use_zero=(mode == "band-zero"))
            dev = self.model.atlas_dyn.device
            sched.dev_items = torch.from_numpy(np.ascontiguousarray(sched.items)).to(dev)
            sched.dev_init = torch.from_numpy(sched.counter_init).to(dev)
            if len(self._sched_cache) >= 8:
                self._sched_cache.pop(next(iter(self._sched_cache)))
            self._sched_cache[key] = sched
        self.last_schedule = sched
        return sched

    def band_rows(self, h, cfg, fit=True, pt_frames=None):
        """(ya, yb): the pixel rows of the target video this rank needs when the loss is sharded by row bands
        (`loss_shard="rows"`): a loader may hand `step()` exactly `res[..., ya:yb, :]`."""
        p, s = int(cfg["patch_size"]), int(cfg["stride"])
        hf = ops._fit(h, p, s, "patch_height") if fit else h
        ho = (hf - p) // s + 1
        if self.world == 1 or self.loss_shard != "rows" or ho < self.world:
            return 0, h
        b = band_layout(h, p, s, ho, self.world)[self.rank]
        return b["ya"], b["yb"]

    def _loss_band_sharded(self, h, w, T, pad, rgb_pad, res0, res_u8, res_ready, cfg, lossobj, gain, sums, grad_rgb, dg):
        """The looping loss sharded by pixel-row bands (SURVEY §8(e) "alternative"): all-to-all of the rendered frames
        into row bands, gain / NN search / vote on the band, all-to-all of dL/drgb back to the frame owners.  Fills
        grad_rgb[t0:t1] (pad adjoint included) and sums[4]."""
        import torch.distributed as dist
        t0, t1 = self.t0, self.t1
        p, s = dg.p, dg.s
        bands = band_layout(h, p, s, dg.ho, self.world)
        me = bands[self.rank]
        ya, yb, hb = me["ya"], me["yb"], me["yb"] - me["ya"]
        F_ = res0.shape[0]
        # ---- target band: bytes -> float on the band only; a full-height target is cropped here
        if res_ready is not None:
            torch.cuda.current_stream().wait_event(res_ready)
        src = res_u8 if res_u8 is not None else res0
        if src.shape[-2] == h and hb != h:
            src = src[:, :, ya:yb]
        elif src.shape[-2] != hb:
            raise Vl3dError(f"target video has {src.shape[-2]} rows; expected the full {h} or this rank's band of {hb}")
        if res_u8 is not None:
            with self._timed("target_u8_to_float"):
                y_band = ops.u8_to_unit(src, out=self._get("y_band", (F_, 3, hb, w), torch.float32))
        else:
            y_band = src if src.is_contiguous() else self._get("y_band", (F_, 3, hb, w), torch.float32).copy_(src)
        # ---- rendered frames -> row bands
        peer = self._peer_exchange(T, pad, h, w, bands, (dg.ho, dg.wo, dg.n1))
        if peer is not None:
            x_band = peer.x_band(self.rank)
            with self._timed("frames_to_bands"):
                peer.frames_to_bands(rgb_pad[t0:t1])                 # NVLink stores into every rank's band (+ loop pad)
        else:
            x_band = self._get("x_band", (T + pad, 3, hb, w), torch.float32)
            with self._timed("frames_to_bands"):
                self._buf["a2a_send"] = frames_to_bands(rgb_pad[t0:t1], x_band, self.bounds, bands, self.rank, self.group,
                                                        self._buf.get("a2a_send"))
                if pad:
                    x_band[T:T + pad].copy_(x_band[:pad])            # loop pad (MPV.py:490-492)
        # ---- scale-invariant gain: partial log-sum over the owned rows, one all-reduce of a double
        xscale = None
        own = (me["own0"] - ya, me["own1"] - ya)
        if self.model.args.scale_invariant:
            with self._timed("scale_invariant"):
                part = self._get("scale_part", (ops._lib.load().vl3d_scale_partials(),), torch.float64)
                lsum = ops.scale_log_sum(x_band, T, y_band, own, part, self._get("scale_lsum", (1,), torch.float64))
                dist.all_reduce(lsum, group=self.group)
                xscale = ops.scale_finish(lsum, 3 * h * w, self._get("xscale", (1,), torch.float32))
        # ---- NN search on the band; the index map of the whole image is assembled in place
        desc = ops.make_loss_desc(x_band.shape, (x_band.stride(0), x_band.stride(1), x_band.stride(2)), y_band.shape,
                                  (y_band.stride(0), y_band.stride(1), y_band.stride(2)), p, dg.pt, s, dg.st,
                                  dg.alpha if dg.use_alpha else 1e10, fit=lossobj.fit)
        if desc.ho != me["pr1"] - (me["pr0"] - me["halo"]) or desc.wo != dg.wo or desc.n1 != dg.n1:
            raise Vl3dError("band layout does not match the loss descriptor")
        rows = [b["pr0"] for b in bands] + [dg.ho]
        if peer is not None:
            nn = self._buf["nn"] = peer.nn(self.rank)
        else:
            nn = self._get("nn", (dg.ho, dg.wo, dg.n1), torch.int32)
            if not rows_equal(rows):
                nn.zero_()
        nn_band = nn[me["pr0"] - me["halo"]:me["pr1"]]
        x_scaled = self._get("xb_scaled", tuple(x_band.shape), torch.float32)
        with self._timed("patchnn_search"):
            ops.patchnn_search(desc, x_band, xscale, y_band, nn_out=nn_band, rows=(me["halo"], desc.ho), scaled_ws=x_scaled)
        with self._timed("exchange_nn"):                            # the vote needs the halo rows of the ranks above
            if peer is not None:
                peer.share_nn(rows, everything=self.gather_nn)
            else:
                exchange_row_bands(nn, rows, self.rank, self.group)
        # ---- votes, robust loss and its gradient for the owned rows of every frame
        g_band = self._get("g_band", (T + pad, 3, hb, w), torch.float32)
        vote_part = self._get("vote_part_band", (ops._lib.load().vl3d_vote_partials(T + pad, hb, w),), torch.float64)
        lo = self._get("loss_out0", (1,), torch.float32)
        with self._timed("vote_loss"):
            ops.vote_loss(desc, x_band, xscale, y_band, nn_band, cfg.get("rou", 0), cfg.get("scaling", 0.2), gain,
                          (T + pad, hb, w), grad_out=g_band, partials=vote_part, loss_out=lo, rows=own,
                          n_total=3 * dg.t * dg.h * dg.w)
            sums[4] += lo[0]
        # ---- dL/drgb back to the frame owners (adjoint of the loop pad first)
        with self._timed("bands_to_frames"):
            if peer is not None:
                peer.bands_to_frames(g_band)                        # NVLink stores into the frame owners' buffers
                grad_rgb[t0:t1].copy_(peer.grad_frames(self.rank))
            else:
                if pad:
                    g_band[:pad] += g_band[T:T + pad]
                self._buf["a2a_send2"], self._buf["a2a_recv2"] = bands_to_frames(
                    g_band, grad_rgb[t0:t1], self.bounds, bands, self.rank, self.group, self._buf.get("a2a_send2"),
                    self._buf.get("a2a_recv2"))

    def _peer_exchange(self, T, pad, h, w, bands, nn_shape):
        """PeerExchange laid out for this problem, or None when the exchanges go through NCCL."""
        if self.exchange == "nccl":
            return None
        if self._peer is None:
            try:
                self._peer = PeerExchange(self.group, self.model.atlas_dyn.device)
                self._peer.layout(T, pad, h, w, bands, self.bounds, nn_shape)
            except Exception as e:                                  # no symmetric memory on this system / build
                if self.exchange == "p2p":
                    raise
                import warnings
                warnings.warn(f"peer-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-to-alls")
                self.exchange, self._peer = "nccl", None
                return None
        return self._peer.layout(T, pad, h, w, bands, self.bounds, nn_shape)

    def _adam(self, name, p, g, lr):
        st = self._state.get(name)
        if st is None:
            st = (torch.zeros_like(p), torch.zeros_like(p))
            self._state[name] = st
        ops.adam_step(p, g, st[0], st[1], self.t, lr, self.betas[0], self.betas[1], self.eps)

    @torch.no_grad()
    def step(self, h, w, tar_extrin, tar_intrin, res, losscfg, lr, optimise=True, res_ready=None):
        """`res_ready`: optional CUDA event after which `res` is valid (e.g. recorded behind an asynchronous
        host-to-device copy on another stream); the render does not need the target video, so the wait is
        placed right before the looping loss and the copy overlaps the render."""
        m = self.model
        args = m.args
        dev = m.atlas_dyn.device
        dist = None
        if self.world > 1:
            import torch.distributed as dist
        atlas_dyn, atlas = m._texels()
        cfg = dict(losscfg)
        loss_name = cfg.pop("loss_name")
        gain = float(cfg.pop("loss_gain", 1.0))
        lossobj = m.losses[loss_name]
        if not isinstance(lossobj, (Patch3DGPNNLowMemLoss, Patch3DGPNNDirectLoss)):
            raise NotImplementedError(f"FusedLoopStep supports the gpnn / gpnn_lm losses, not {loss_name!r}")
        _check_dist(cfg)
        if (len(getattr(args, "bg_color", "")) > 0
                or any(getattr(args, f"{k}_loss_weight", 0) > 0 for k in ("sparsity", "density", "d_smooth"))):
            raise NotImplementedError("FusedLoopStep covers the terms the shipped stage-2 configs use; bg_color / sparsity / "
                                      "density / d_smooth run on the autograd path (MPMeshVid.forward + loss.backward(), "
                                      "i.e. make_run_iter)")
        T, t0, t1 = self.T, self.t0, self.t1
        Tl = t1 - t0
        pad = m.swd_patcht_size - 1 if m.isloop else 0
        if pad > T:
            raise ValueError("loop pad longer than the video")
        res0 = res[0] if res.dim() == 5 else res
        res_u8 = res0 if res0.dtype == torch.uint8 else None       # bytes from the loader: converted after `res_ready`
        if res_u8 is None and not res0.is_contiguous():
            res0 = res0.contiguous()
        # pose / intrinsics are host data (the view descriptor is built on the host); a device tensor costs a sync here
        ext = _host64(tar_extrin).reshape(4, 4) @ m.ref_extrin_inv_host()
        view = m.make_view(h, w, ext, _host64(tar_intrin))
        pack = m._pack
        wr, wa = args.rgb_smooth_loss_weight, args.a_smooth_loss_weight
        smooth = wr > 0 or wa > 0
        dyn_local = atlas_dyn.data if self.local_model else atlas_dyn.data[t0:t1]

        # ---- render.  With a backward pass coming the regulariser sums are produced there (it exchanges the
        # same neighbour values anyway) and the forward stays a pure render.
        rgb_pad = self._get("rgb_pad", (T + pad, 3, h, w), torch.float32)
        sums = self._get("sums", (5,), torch.float64)             # 4 regulariser sums + the loss partial
        sums.zero_()
        fwd_sums = sums[:4] if (smooth and not optimise) else None
        with self._timed("composite_fwd"):
            if self.world == 1:
                ops.composite_fwd(view, pack, dyn_local, atlas.data, None, T, pad, rgb_out=rgb_pad, smooth_sums=fwd_sums)
            else:
                ops.composite_fwd(view, pack, dyn_local, atlas.data, None, Tl, 0, rgb_out=rgb_pad[t0:t1],
                                  smooth_sums=fwd_sums)
        band_mode = False
        if self.world > 1 and self.loss_shard == "rows":
            dg = ops.make_loss_desc(rgb_pad.shape, (3 * h * w, h * w, w), (res0.shape[0], 3, h, w), (3 * h * w, h * w, w),
                                    cfg["patch_size"], cfg["patcht_size"], cfg["stride"], cfg["stridet"],
                                    cfg.get("alpha", 1e10), fit=lossobj.fit)
            band_mode = dg.ho >= self.world                         # (fewer patch rows than ranks: frame sharding)
        if band_mode:
            grad_rgb = self._get("grad_rgb", (T + pad, 3, h, w), torch.float32)
            self._loss_band_sharded(h, w, T, pad, rgb_pad, res0, res_u8, res_ready, cfg, lossobj, gain, sums, grad_rgb, dg)
            desc = dg
        gather = None
        if self.world > 1 and not band_mode:
            # asynchronous: the target-frame sums of the scale-invariant gain (which do not need the rendered video)
            # run on the compute stream while NVLink moves the frames; waited for right before the gain is evaluated
            with self._timed("allgather_rgb_issue"):
                gather = gather_frames(rgb_pad, self.bounds, self.rank, T, self.group, async_op=True)

        def finish_gather():
            nonlocal gather
            if gather is not None:
                gather.wait()
                gather = None
                if pad:
                    rgb_pad[T:T + pad].copy_(rgb_pad[:pad])          # loop pad (MPV.py:490-492)

        # ---- looping loss (single GPU, or frame-sharded: every rank holds the whole rendered and target video)
        if not band_mode:
            if res_ready is not None:
                torch.cuda.current_stream().wait_event(res_ready)
            if res_u8 is not None:                                      # `vid / 255` of the dataset (train_3dvid.py:54), on device
                with self._timed("target_u8_to_float"):
                    res0 = ops.u8_to_unit(res_u8, out=self._get("res_f32", tuple(res_u8.shape), torch.float32))
            xscale = None
            if args.scale_invariant:
                with self._timed("scale_invariant"):
                    out = self._get("xscale", (1,), torch.float32)
                    part = self._get("scale_part", (ops._lib.load().vl3d_scale_partials(),), torch.float64)
                    if self.world == 1:
                        xscale = ops.scale_invariant(rgb_pad, T, res0, out=out, partials=part)
                    else:
                        # the mean over the F target frames is the expensive part (the whole target video is read):
                        # every rank sums its block of frames, one all-reduce of the (3,h,w) sums
                        fb = partition(res0.shape[0], self.world)
                        rsum = ops.frame_sum(res0[fb[self.rank]:fb[self.rank + 1]], out=self._get("res_sum", (3, h, w), torch.float32))
                        dist.all_reduce(rsum, group=self.group)
                        finish_gather()
                        xscale = ops.scale_invariant_presum(rgb_pad, T, rsum, res0.shape[0], out=out, partials=part)
            finish_gather()
            desc = ops.make_loss_desc(rgb_pad.shape, (rgb_pad.stride(0), rgb_pad.stride(1), rgb_pad.stride(2)), res0.shape,
                                      (res0.stride(0), res0.stride(1), res0.stride(2)), cfg["patch_size"],
                                      cfg["patcht_size"], cfg["stride"], cfg["stridet"], cfg.get("alpha", 1e10),
                                      fit=lossobj.fit)
            nn = self._get("nn", (desc.ho, desc.wo, desc.n1), torch.int32)
            x_scaled = self._get("x_scaled", tuple(rgb_pad.shape), torch.float32)
            if self.world == 1:
                with self._timed("patchnn_search"):
                    ops.patchnn_search(desc, rgb_pad, xscale, res0, nn_out=nn, scaled_ws=x_scaled)
            else:
                # patch positions are independent (utils_vid.py:211-215): each rank searches a band of patch rows,
                # then the int32 index map is summed across ranks (disjoint rows, zeros elsewhere)
                rows = partition(desc.ho, self.world)
                if not rows_equal(rows):
                    nn.zero_()
                with self._timed("patchnn_search"):
                    ops.patchnn_search(desc, rgb_pad, xscale, res0, nn_out=nn, rows=(rows[self.rank], rows[self.rank + 1]),
                                       scaled_ws=x_scaled)
                with self._timed("exchange_nn"):
                    exchange_row_bands(nn, rows, self.rank, self.group)
            grad_rgb = self._get("grad_rgb", (T + pad, 3, h, w), torch.float32)
            n_part = ops._lib.load().vl3d_vote_partials(T + pad, h, w)
            vote_part = self._get("vote_part", (n_part,), torch.float64)
            with self._timed("vote_loss"):
                ranges = [(0, T + pad)] if self.world == 1 else owned_frame_ranges(self.bounds, self.rank, T, pad)
                for i, fr in enumerate(ranges):
                    lo = self._get(f"loss_out{i}", (1,), torch.float32)
                    ops.vote_loss(desc, rgb_pad, xscale, res0, nn, cfg.get("rou", 0), cfg.get("scaling", 0.2), gain,
                                  (T + pad, h, w), grad_out=grad_rgb, partials=vote_part, loss_out=lo, frames=fr)
                    sums[4] += lo[0]

        # d total / d smooth_sums (host constants): MPV.py:517-531 with K cancelled, train_3dvid.py:230-240
        nx = max(T * h * (w - 1), 1) * m.mpi_d
        ny = max(T * (h - 1) * w, 1) * m.mpi_d
        w_smooth = None
        if smooth:
            key = ("w_smooth", T, h, w, gain, wr, wa)
            w_smooth = self._buf.get(key)
            if w_smooth is None:
                w_smooth = torch.tensor([wr * gain / (3 * nx), wr * gain / (3 * ny), wa * gain / nx, wa * gain / ny],
                                        dtype=torch.float32, device=dev)
                self._buf[key] = w_smooth

        def assemble():
            if self.world > 1:
                dist.all_reduce(sums, group=self.group)              # 5 doubles: regulariser sums + loss partials
            out = {"swd": (sums[4] * gain).float()}
            total = out["swd"]
            if wr > 0:
                out["rgb_smooth"] = ((sums[0] / (3 * nx) + sums[1] / (3 * ny)) * gain).float()
                total = total + out["rgb_smooth"] * wr
            if wa > 0:
                out["a_smooth"] = ((sums[2] / nx + sums[3] / ny) * gain).float()
                total = total + out["a_smooth"] * wa
            out["loss"] = total
            return out

        if not optimise:
            return assemble()

        # ---- backward + Adam on the owned frames
        g_sta = self._like("g_sta", atlas.data)
        if pack.n_static > 0:
            with self._timed("grad_sta_zero"):
                g_sta.zero_()
        bwd_sums = sums[:4] if smooth else None
        # adjoint of the loop pad for the frames we own, so the backward can run per frame chunk with pad = 0
        if pad and t0 < pad and not band_mode:                      # (band mode: folded before the exchange)
            n = min(t1, pad) - t0
            grad_rgb[t0:t0 + n] += grad_rgb[T + t0:T + t0 + n]
        st = self._state.get("atlas_dyn")
        if st is not None and st[0].shape != dyn_local.shape:      # lod() changed the atlas: fresh optimiser state
            self.reset()
            st = None
        if st is None:
            st = (torch.zeros_like(dyn_local), torch.zeros_like(dyn_local))
            self._state["atlas_dyn"] = st
        self.t += 1
        mode = self.fused
        if mode == "auto":
            mode = "band" if (pack.rect_planes and Tl // 2 <= 6 and h * w >= 512 * 512) else "generic"
        if (mode.startswith("band") or mode == "own") and not pack.rect_planes:
            mode = "generic"
        if mode == "own" and w_smooth is None:                      # the owner path rides on the regulariser tiling
            mode = "generic"
        Te = Tl // 2 * 2 if mode != "off" else 0                   # frames handled by the fused kernel (chunks of 2)
        if Te > 0:
            sched = self._schedule_for(mode, view, pack, h, w, smooth, dyn_local.shape[-2:])
            n_rounds = Te // 2 + (1 if sched.extra_round else 0)
            g_dyn = self._buf.get("g_dyn")
            if g_dyn is None or g_dyn.shape != dyn_local.shape or tuple(g_dyn.stride()) != tuple(dyn_local.stride()):
                g_dyn = None
                if self.grad_compress:                              # zeros cross HBM compressed (ops.compressible_zeros_like)
                    g_dyn, self.grad_compressed = ops.compressible_zeros_like(dyn_local)
                if g_dyn is None:
                    g_dyn = torch.zeros_like(dyn_local)            # all-zero between steps (the kernel keeps it so)
                self._buf["g_dyn"] = g_dyn
            state = self._get("fused_state", (16 + n_rounds * sched.n_counters,), torch.int32)
            state[:16].zero_()                                      # [0] = queue head
            state[16:].view(n_rounds, sched.n_counters).copy_(sched.dev_init)   # every round's counters
            with self._timed("fused_bwd_adam"):
                if mode == "own":
                    scratch = self._buf.get("own_scratch")
                    if scratch is None:
                        scratch = self._buf["own_scratch"] = ops.fused_own_scratch(dyn_local.device)
                    table = self._buf.get(("own_table", h, w))
                    if table is None:
                        table = self._buf[("own_table", h, w)] = ops.fused_own_table(dyn_local.device, h, w)
                    ops.fused_bwd_adam_own(view, pack, dyn_local[:Te], atlas.data, Te, grad_rgb[t0:t0 + Te], rgb_pad[t0:t0 + Te],
                                           w_smooth, bwd_sums, g_dyn[:Te], g_sta, st[0][:Te], st[1][:Te], self.t, lr,
                                           self.betas[0], self.betas[1], self.eps, sched.dev_items, sched.n_items, n_rounds,
                                           state, sched.n_counters, sched.own, table, scratch,
                                           ctas_per_sm=self.fused_opts.get("ctas_per_sm", 0))
                else:
                    ops.fused_bwd_adam(view, pack, dyn_local[:Te], atlas.data, Te, grad_rgb[t0:t0 + Te], rgb_pad[t0:t0 + Te],
                                       w_smooth, bwd_sums, g_dyn[:Te], g_sta, st[0][:Te], st[1][:Te], self.t, lr,
                                       self.betas[0], self.betas[1], self.eps, sched.dev_items, sched.n_items, n_rounds,
                                       state, sched.n_counters, ctas_per_sm=self.fused_opts.get("ctas_per_sm", 0))
        if Te < Tl:
            # separate kernels: everything when fused == "off", else the odd last frame
            g_dyn = self._buf.get("g_dyn")
            if g_dyn is None or g_dyn.shape != dyn_local.shape or tuple(g_dyn.stride()) != tuple(dyn_local.stride()):
                g_dyn = torch.zeros_like(dyn_local)
                self._buf["g_dyn"] = g_dyn
            with self._timed("grad_zero"):
                g_dyn[Te:].zero_()
            with self._timed("composite_bwd"):
                ops.composite_bwd(view, pack, dyn_local[Te:], atlas.data, None, Tl - Te, 0, grad_rgb[t0 + Te:t1],
                                  rgb_pad[t0 + Te:t1], w_smooth, g_dyn[Te:], g_sta, smooth_sums=bwd_sums)
            with self._timed("adam"):
                ops.adam_step(dyn_local[Te:], g_dyn[Te:], st[0][Te:], st[1][Te:], self.t, lr, self.betas[0],
                              self.betas[1], self.eps)
            if Te > 0 and not mode.endswith("zero"):
                g_dyn[Te:].zero_()                                  # keep the fused kernel's all-zero invariant
        if pack.n_static > 0:
            if self.world > 1:
                with self._timed("allreduce_static_grad"):
                    dist.all_reduce(g_sta, group=self.group)         # the one gradient all-reduce
            with self._timed("adam_static"):
                self._adam("atlas", atlas.data, g_sta, lr)
        return assemble()


def _host64(a):
    """float64 numpy copy of a pose / intrinsics argument (host tensors and arrays: no device sync)."""
    if torch.is_tensor(a):
        return a.detach().double().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


class _Timed:
    """CUDA-event bracket on the current stream (the stream the kernels are launched on)."""

    def __init__(self, store, name):
        self.store, self.name = store, name

    def __enter__(self):
        if self.store is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.store is not None:
            self.b.record()
            self.store.setdefault(self.name, []).append((self.a, self.b))
        return False
