"""Per-term gradient comparison of the CULLED stage-1 model (CUDA) against the oracle (diagnostic, not a test)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import mpv_oracle as MO
from util import load_golden
from videoloop3d_b200 import MPMesh, default_args_stage1
dev = torch.device("cuda:0")
g = load_golden("stage1_sparsify")
H, W, D, hv, wv = (int(g[k]) for k in ("H", "W", "D", "hv", "wv"))
args = default_args_stage1(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpi_h_scale=1.0, mpi_w_scale=1.0, d_smooth_loss_weight=0.1)
f = 0.8 * W
m = MPMesh(args, H, W, np.eye(4, dtype=np.float32), np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32), 1.0, 10.0)
m.atlas.data = torch.as_tensor(g["atlas0"]).clone(); m.atlas_mask.data = torch.as_tensor(g["atlas_mask0"]).clone()
m = m.to(dev); m.sparsify_faces(erode_num=int(g["erode_num"]), alpha_thresh=float(g["alpha_thresh"]))
ext, intr = torch.as_tensor(g["tar_extrin"]), torch.as_tensor(g["tar_intrin"])
sd = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in m.state_dict().items()}
st = MO.MPVState.from_state_dict(sd, D, hv, wv)
a = st.atlas.double().requires_grad_(True); ad = st.atlas_dyn.double().requires_grad_(True)
rgbl_o, extra_o, var_o = MO.forward_stage1(st, H, W, ext, intr, 1.0, 10.0, atlas=a, atlas_dyn=ad, l_smooth=False)
gen = torch.Generator().manual_seed(5)
g_up = torch.rand(rgbl_o.shape, generator=gen, dtype=torch.float64) - 0.4
m.train()
terms = dict(rgb=lambda r, e: (r * (g_up if r.dtype == torch.float64 else g_up.to(dev).float())).mean(),
             **{k: (lambda r, e, k=k: e[k].mean()) for k in ("sparsity", "rgb_smooth", "a_smooth", "density", "d_smooth")})
for name, fn in terms.items():
    go, gdo = torch.autograd.grad(fn(rgbl_o, extra_o), (a, ad), retain_graph=True, allow_unused=True)
    rgbl, extra = m(H, W, ext.to(dev), intr.to(dev))
    gc, gdc = torch.autograd.grad(fn(rgbl, extra), (m.atlas, m.atlas_dyn), allow_unused=True)
    for pn, o, c in (("atlas", go, gc), ("atlas_dyn", gdo, gdc)):
        d = (c.cpu().double() - o).abs()
        idx = np.unravel_index(int(d.argmax()), d.shape)
        print(f"{name:10s} {pn:9s} max|ref| {float(o.abs().max()):.3e} max err {float(d.max()):.3e} n(err>5e-4 max) {int((d > 5e-4 * o.abs().max()).sum())} "
              f"at {tuple(int(i) for i in idx)} ref {float(o[idx]):.3e} got {float(c.cpu()[idx]):.3e}")
