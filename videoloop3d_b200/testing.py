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
