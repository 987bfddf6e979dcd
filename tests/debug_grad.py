"""GPU debug helper: where do CUDA and oracle gradients differ?  (development tool)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from oracle import mpv_oracle as MO
from test_gpu_parity import CASES, _make_case, state_tensors
from videoloop3d_b200.testing import model_from_tensors

dev = torch.device("cuda:0")
for case in CASES:
    st, ext, intr, res, cfg = _make_case(case)
    H, W = case["H"], case["W"]
    for (wr, wa, tag) in [(0.0, 0.0, "no-smooth"), (0.2, 0.2, "smooth")]:
        a = st.atlas.double().requires_grad_(True); ad = st.atlas_dyn.double().requires_grad_(True)
        extra, aux = MO.forward_train(st, H, W, ext, intr, res, cfg, swd_patcht_size=case["pt"], atlas=a, atlas_dyn=ad,
                                      rgb_smooth=wr > 0, a_smooth=wa > 0)
        MO.total_loss(extra, wr, wa).backward()
        m = model_from_tensors(state_tensors(st), H, W, dev, swd_patcht_size=case["pt"], rgb_smooth_loss_weight=wr,
                               a_smooth_loss_weight=wa)
        m.train()
        batched = {k: ([v] if isinstance(v, str) else torch.tensor([v])) for k, v in cfg.items()}
        _, ex = m(H, W, ext.to(dev), intr.to(dev), res=res.to(dev), losscfg=batched)
        loss = ex["swd"].mean()
        if wr > 0:
            loss = loss + wr * ex["rgb_smooth"].mean() + wa * ex["a_smooth"].mean()
        loss.backward()
        gc = m.atlas_dyn.grad.cpu().double(); go = ad.grad
        err = (gc - go).abs()
        idx = np.unravel_index(int(err.argmax()), err.shape)
        big = (err > 1e-4 * go.abs().max()).sum().item()
        print(f"{case['kind']}-D{case['D']}-p{case['p']} {tag}: relerr {float(err.max()/go.abs().max()):.2e} at {idx} "
              f"cuda {float(gc[idx]):.4e} oracle {float(go[idx]):.4e} gmax {float(go.abs().max()):.3e} n>1e-4: {big} / {int((go!=0).sum())}")

# hit-mask comparison for the first case
case = CASES[0]
st, ext, intr, res, cfg = _make_case(case)
H, W = case["H"], case["W"]
m = model_from_tensors(state_tensors(st), H, W, dev, swd_patcht_size=case["pt"])
m.eval()
with torch.no_grad():
    rgb, var = m.render(H, W, ext.to(dev) @ m.ref_extrin[None].inverse(), intr.to(dev), [0])
    hits_c = (var["pix_to_face"][0] >= 0).sum(-1).cpu()
geo = MO.geometry(st, H, W, ext, intr)
hits_o = geo["hit"].sum(1).reshape(H, W)
d = (hits_c != hits_o).nonzero()
print("hit-count mismatches:", d.tolist())
for (r, c) in d.tolist():
    pix = r * W + c
    print(" pixel", r, c, "oracle hit", geo["hit"][pix].int().tolist(), "cuda hits", int(hits_c[r, c]))
    # oracle grid coords
    D, hv, wv = st.mpi_d, st.hv, st.wv
print("K cuda", var["mpi"].shape[-2])
