"""Per-term gradient comparison of MPMesh (CUDA) against the oracle for one view (diagnostic, not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import mpv_oracle as MO
from test_gpu_composite_sweep import VIEWS, _rot
from videoloop3d_b200 import MPMesh, default_args_stage1

vname = sys.argv[1] if len(sys.argv) > 1 else "roll"
dev = torch.device("cuda:0")
H, W, D, hv, wv = 37, 70, 8, 6, 9
seed = 21 + sorted(VIEWS).index(vname)
st, atlas_mask = MO.stage1_state(H, W, D, hv, wv, 2, 1.0, 10.0, 1.6, 1.6, seed=seed)
v = VIEWS[vname]
ext = torch.eye(4); ext[:3, :3] = _rot(*v["rot"]); ext[:3, 3] = torch.tensor(v["trans"])
f = 0.8 * W * v["fmul"]
intr = torch.tensor([[f, 0, W / 2 + 0.37], [0, f, H / 2 - 0.21], [0, 0, 1.]])
weights = dict(sparsity=0.3, rgb_smooth=0.2, a_smooth=0.5, density=0.2, d_smooth=0.4, l_smooth=0.1)
a = st.atlas.double().requires_grad_(True); am = atlas_mask.double().requires_grad_(True)
rgbl_o, extra_o, var_o = MO.forward_stage1(st, H, W, ext[None], intr[None], 1.0, 10.0, edge_scale=0.5, atlas=a, atlas_mask=am)
args = default_args_stage1(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=2, mpi_h_scale=1.6, mpi_w_scale=1.6, edge_scale=0.5,
                           **{k + "_loss_weight": w for k, w in weights.items()})
fr = 0.8 * W
m = MPMesh(args, H, W, np.eye(4, dtype=np.float32), np.array([[fr, 0, W / 2], [0, fr, H / 2], [0, 0, 1]], dtype=np.float32), 1.0, 10.0)
m.atlas.data, m.atlas_mask.data = st.atlas.clone(), atlas_mask.clone()
m = m.to(dev).train()
gen = torch.Generator().manual_seed(seed)
g_up = torch.rand(rgbl_o.shape, generator=gen, dtype=torch.float64) - 0.4
terms = dict(rgb=lambda r, e: (r * (g_up if r.dtype == torch.float64 else g_up.to(dev).float())).mean(), **{k: (lambda r, e, k=k: e[k].mean()) for k in weights})
for name, fn in terms.items():
    go, gmo = torch.autograd.grad(fn(rgbl_o, extra_o), (a, am), retain_graph=True, allow_unused=True)
    rgbl, extra = m(H, W, ext[None].to(dev), intr[None].to(dev))
    gc, gmc = torch.autograd.grad(fn(rgbl, extra), (m.atlas, m.atlas_mask), allow_unused=True)
    for pn, o, c in (("atlas", go, gc), ("mask", gmo, gmc)):
        if o is None:
            print(name, pn, "oracle None; cuda", None if c is None else float(c.abs().max()))
            continue
        d = (c.cpu().double() - o).abs()
        idx = np.unravel_index(int(d.argmax()), d.shape)
        print(f"{name:10s} {pn:6s} max|ref| {float(o.abs().max()):.3e} max err {float(d.max()):.3e} rel {float(d.max() / o.abs().max()):.2e} "
              f"n(err>1e-3 max) {int((d > 1e-3 * o.abs().max()).sum())} at {idx} ref {float(o[idx]):.3e} got {float(c.cpu()[idx]):.3e}")
    val_o = float(fn(rgbl_o, extra_o)); val_c = float(fn(rgbl, extra))
    print(f"   value oracle {val_o:.8f} cuda {val_c:.8f}")
