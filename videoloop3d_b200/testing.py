"""Helpers shared by tests / smoke / bench: build an `MPMeshVid` around given tensors."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .mpv import MPMeshVid
from .train_step import default_args


def model_from_tensors(t, H, W, device, args=None, **arg_overrides):
    """`t`: dict with verts, planedepth, faces(_dyn), uvs(_dyn), uvfaces(_dyn), atlas(_dyn), ref_extrin,
    ref_intrin, mpi_d, hv, wv (numpy or torch).  Returns an MPMeshVid on `device` holding exactly them."""
    g = lambda k: torch.as_tensor(np.asarray(t[k])) if not torch.is_tensor(t[k]) else t[k].detach()
    D, hv, wv = int(t["mpi_d"]), int(t["hv"]), int(t["wv"])
    T = g("atlas_dyn").shape[0]
    if args is None:
        args = default_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=1, mpv_frm_num=1,
                            mpi_h_scale=0.1, mpi_w_scale=0.1, **arg_overrides)
    m = MPMeshVid(args, H, W, np.eye(4, dtype=np.float32), np.eye(3, dtype=np.float32), 1.0, 10.0)
    m._verts.data = g("verts").float()
    m.planedepth.data = g("planedepth").float()
    m.ref_extrin.data = g("ref_extrin").float()
    m.ref_intrin.data = g("ref_intrin").float()
    m.uvs.data = g("uvs").float()
    m.uvs_dyn.data = g("uvs_dyn").float()
    m.uvfaces = g("uvfaces").long()
    m.uvfaces_dyn = g("uvfaces_dyn").long()
    m.faces = g("faces").long()
    m.faces_dyn = g("faces_dyn").long()
    m.register_parameter("atlas", nn.Parameter(g("atlas").float().clone()))
    m.register_parameter("atlas_dyn", nn.Parameter(g("atlas_dyn").float().clone()))
    m.frm_num = T
    m.is_sparse, m.has_dyn = True, True
    m = m.to(device)
    m.atlas.data = ops.as_texels(m.atlas.data)
    m.atlas_dyn.data = ops.as_texels(m.atlas_dyn.data)
    m.invalidate_geometry()
    return m


def cull_to_tiles(m, tile, occupancy, dyn_frac, frames, seed=0, device=None, first_frame=0, alpha_mean=-1.0):
    """Turn a dense `MPMeshVid` into a tile-culled one with random content (synthetic stand-in for a stage-1 result
    that went through `MPI.sparsify_faces`, MPI.py:289-442): every quad is kept with probability `occupancy`, a kept
    quad is dynamic with probability `dyn_frac`; each kept quad owns a private `tile` x `tile` texel square, squares
    are packed row-major into a static and a dynamic atlas, `uvs(_dyn)` hold four private corners per quad
    (u = x / (W - 1) * 2 - 1), `faces(_dyn)` keep indexing the shared vertex grid.  The dynamic atlas gets `frames`
    frames; frame t's texels depend on (seed, first_frame + t) only, the static atlas on `seed` only."""
    device = torch.device(device) if device is not None else m.atlas_dyn.device
    D, hv, wv = m.mpi_d, m.mpi_h_verts, m.mpi_w_verts
    nq = D * (hv - 1) * (wv - 1)
    g = torch.Generator().manual_seed(seed)
    keep = torch.rand(nq, generator=g) < occupancy
    dyn = torch.rand(nq, generator=g) < dyn_frac
    quad_faces = m.faces_dyn.detach().cpu().reshape(nq, 2, 3)
    corner = torch.tensor([[0, 1, 3], [3, 2, 0]])

    def layout(mask):
        ids = torch.nonzero(mask).reshape(-1)
        n = len(ids)
        if n == 0:
            return (1, 1), torch.zeros(0, 2), torch.zeros(0, 3, dtype=torch.long), torch.zeros(0, 3, dtype=torch.long), (1, 1)
        nh = max(1, int(round((n / 2) ** 0.5)))
        nw = (n + nh - 1) // nh
        ah, aw = nh * tile, nw * tile
        k = torch.arange(n)
        x0, y0 = (k % nw) * tile, (k // nw) * tile
        xs = torch.stack([x0, x0 + tile - 1, x0, x0 + tile - 1], 1).double()
        ys = torch.stack([y0, y0, y0 + tile - 1, y0 + tile - 1], 1).double()
        uv = torch.stack([xs / (aw - 1) * 2 - 1, ys / (ah - 1) * 2 - 1], -1).reshape(-1, 2).float()
        uvf = (k[:, None, None] * 4 + corner[None]).reshape(-1, 3)
        return (ah, aw), uv, uvf, quad_faces[ids].reshape(-1, 3), (nh, nw)

    (sh, sw), uvs, uvfaces, faces, sgrid = layout(keep & ~dyn)
    (dh, dw), uvs_dyn, uvfaces_dyn, faces_dyn, dgrid = layout(keep & dyn)
    gd = torch.Generator(device=device)
    gd.manual_seed(seed * 100003 + 7)
    atlas = torch.empty((1, sh, sw, 4), device=device).normal_(generator=gd)
    atlas[..., 3] += alpha_mean
    atlas_dyn = torch.empty((frames, dh, dw, 4), device=device)
    for t in range(frames):
        gd.manual_seed(seed * 100003 + 1000 + first_frame + t)
        atlas_dyn[t].normal_(generator=gd)
    atlas_dyn[..., 3] += alpha_mean
    m.uvs.data = uvs.to(device)
    m.uvs_dyn.data = uvs_dyn.to(device)
    m.uvfaces = uvfaces.long().to(device)
    m.uvfaces_dyn = uvfaces_dyn.long().to(device)
    m.faces = faces.long().to(device)
    m.faces_dyn = faces_dyn.long().to(device)
    m.atlas.data = atlas.permute(0, 3, 1, 2)                         # logical (1,4,H,W), RGBA-interleaved memory
    m.atlas_dyn.data = atlas_dyn.permute(0, 3, 1, 2)
    m.frm_num = frames
    m.is_sparse, m.has_dyn = True, True
    m.atlas_grid_h, m.atlas_grid_w = sgrid
    m.atlas_full_h, m.atlas_full_w = sh, sw
    m.atlas_grid_dyn_h, m.atlas_grid_dyn_w = dgrid
    m.atlas_full_dyn_h, m.atlas_full_dyn_w = dh, dw
    m.invalidate_geometry()
    return m
