"""Shared helpers for the tests: golden-vector loading and tolerances."""
import os

import numpy as np
import torch

from oracle import mpv_oracle as MO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def state_from_golden(g):
    t = lambda k: torch.as_tensor(g[k])
    return MO.MPVState(verts=t("verts"), planedepth=t("planedepth"), faces=t("faces").long(),
                       faces_dyn=t("faces_dyn").long(), uvs=t("uvs"), uvs_dyn=t("uvs_dyn"),
                       uvfaces=t("uvfaces").long(), uvfaces_dyn=t("uvfaces_dyn").long(), atlas=t("atlas"),
                       atlas_dyn=t("atlas_dyn"), ref_extrin=t("ref_extrin"), ref_intrin=t("ref_intrin"),
                       mpi_d=int(g["mpi_d"]), hv=int(g["hv"]), wv=int(g["wv"]))


def cfg_from_golden(g):
    cfg = {}
    for k, v in g.items():
        if k.startswith("cfg_"):
            v = v.item() if v.ndim == 0 else v
            cfg[k[4:]] = v
    return cfg


def relerr(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def ckpt_dict(g, prefix):
    """A reference state dict stored by oracle/make_golden.py::golden_ckpt under `prefix` ('stage1_', 'sd_', 'sd3_')."""
    out = {}
    for k, v in g.items():
        if k.startswith(prefix):
            name = k[len(prefix):].replace("self_", "self.")
            out[name] = (v.item() if v.ndim == 0 else torch.as_tensor(v)) if name.startswith("self.") else torch.as_tensor(v)
    return out


def ckpt_model(g, device):
    """A fresh stage-2 `MPMeshVid` of the golden_ckpt configuration."""
    from videoloop3d_b200 import MPMeshVid, default_args
    H, W = int(g["H"]), int(g["W"])
    args = default_args(mpi_d=int(g["D"]), mpi_h_verts=int(g["hv"]), mpi_w_verts=int(g["wv"]), atlas_grid_h=2,
                        mpv_frm_num=int(g["T"]), mpi_h_scale=1.3, mpi_w_scale=1.3)
    f = 0.8 * W
    m = MPMeshVid(args, H, W, np.eye(4, dtype=np.float32),
                  np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32), 1.0, 10.0)
    return m.to(device)
