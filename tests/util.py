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
