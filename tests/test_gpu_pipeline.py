"""GPU: the two stages chained on synthetic data, the way the reference's scripts chain them (train_3d.py -> checkpoint ->
train_3dvid.py): a few stage-1 steps on `MPMesh`, tile culling, a few more steps on the culled model, `state_dict`,
`MPMeshVid.init_from_mpi`, a few fused stage-2 steps.  An integration check (shapes, devices, formats, finiteness, the loss
goes down); the numerics of every piece are pinned elsewhere (tests/test_gpu_stage1.py, test_gpu_parity.py, ...)."""
import numpy as np
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu


def test_stage1_to_stage2_pipeline():
    from videoloop3d_b200 import (FusedLoopStep, MPMesh, MPMeshVid, default_args, default_args_stage1, make_run_iter_stage1)
    dev = torch.device("cuda:0")
    g = load_golden("stage1_sparsify")                              # blobs of content on an untouched background
    H, W, D, hv, wv = (int(g[k]) for k in ("H", "W", "D", "hv", "wv"))
    f = 0.8 * W
    intr0 = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    args1 = default_args_stage1(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpi_h_scale=1.0, mpi_w_scale=1.0,
                                add_intrin_noise=False, lrate=0.05)
    m1 = MPMesh(args1, H, W, np.eye(4, dtype=np.float32), intr0, 1.0, 10.0)
    m1.atlas.data = torch.as_tensor(g["atlas0"]).clone()
    m1.atlas_mask.data = torch.as_tensor(g["atlas_mask0"]).clone()
    m1 = m1.to(dev)
    ext, intr = torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"])
    gen = torch.Generator().manual_seed(1)
    data = (0, 0, torch.inverse(ext)[:, :3, :], intr, torch.rand(1, 3, H, W, generator=gen) * 0.5 + 0.25,
            (torch.rand(1, H, W, generator=gen) > 0.5).float())
    run_iter = make_run_iter_stage1(args1, m1, dev)
    opt = m1.get_optimizer()
    losses = [float(run_iter(i, opt, data)) for i in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    info = m1.sparsify_faces(erode_num=1, alpha_thresh=0.05)
    assert m1.has_dyn and info["kept"] >= info["dynamic"] >= 16
    opt = m1.get_optimizer()                                        # train_3d.py:285: a new optimiser over the new tensors
    assert sum(len(gr["params"]) for gr in opt.param_groups) == 5   # uvs, atlas, uvs_dyn, atlas_dyn | _verts
    more = [float(run_iter(6 + i, opt, data)) for i in range(3)]
    assert all(np.isfinite(more)) and more[-1] < more[0] * 1.02, more      # (a fresh Adam's first steps: no strict descent asked)
    sd = m1.state_dict()
    # stage 2 (train_3dvid.py:147, 203-211)
    T = 4
    args2 = default_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpv_frm_num=T, mpi_h_scale=1.0, mpi_w_scale=1.0,
                         swd_patcht_size=3)
    m2 = MPMeshVid(args2, H, W, np.eye(4, dtype=np.float32), intr0, 1.0, 10.0).to(dev)
    m2.init_from_mpi(sd)
    assert tuple(m2.atlas_dyn.shape) == (T,) + tuple(m1.atlas_dyn.shape[1:]) and tuple(m2.atlas.shape) == tuple(m1.atlas.shape)
    assert torch.equal(m2.faces_dyn, m1.faces_dyn) and m2.is_sparse and m2.has_dyn
    # the culled stage-1 frame, replicated over T, renders in stage 2 exactly what stage 1 renders
    m1.eval(); m2.eval()
    with torch.no_grad():
        r1, _ = m1(H, W, ext.to(dev), intr.to(dev))
        r2, _ = m2(H, W, ext.to(dev), intr.to(dev))
    assert float((r2 - r1[:, :3]).abs().max()) < 1e-5
    cfg = dict(loss_name="gpnn_lm", loss_gain=1.0, patch_size=5, patcht_size=3, stride=2, stridet=1, alpha=0.0, rou="-2",
               scaling=0.1, dist_fn="mse", macro_block=65, factor=1)
    res = torch.rand(1, 8, 3, H, W, generator=gen).to(dev)
    before = m2.atlas_dyn.detach().clone()
    step = FusedLoopStep(m2)
    outs = [step.step(H, W, ext, intr, res, cfg, lr=0.01) for _ in range(3)]
    assert all(bool(torch.isfinite(o["loss"])) for o in outs)
    assert float((m2.atlas_dyn.detach() - before).abs().max()) > 1e-3
