"""GPU: the looping-loss kernels across BASELINE config 5's axes (patch 7/11/15, T up to 96, up to 1024
candidates, both alpha modes) plus the odd corners (temporal stride 2, pt = 1, patches that do not overlap,
the un-fitted direct loss, non-contiguous inputs) — each against the CPU oracle on a small spatial extent.
Every kernel variant is hit: strip kernel M = 0..3, 16-byte and 4-byte staging, the 4x8-tile strip kernel with
TMA and with LDGSTS staging, the one-patch-per-CTA fallback (more than 64 query frames), the small-patch kernel."""
import pytest
import torch

from oracle import looploss_oracle as LL

pytestmark = pytest.mark.gpu

CASES = [
    # name,           t,  F,   h,  w,  p, pt, s, st, alpha, rou,   cls
    ("p7_M1_vec",     12, 40,  31, 36, 7, 3, 4, 1, 0.0, "-2", "lm"),
    ("p11_M2_vec",    10, 70,  35, 44, 11, 3, 4, 1, 1e4, "-2", "lm"),
    ("p15_M3_vec",    8,  30,  43, 48, 15, 3, 4, 1, 0.0, "0", "lm"),
    ("p3_s2_M1",      9,  20,  21, 27, 3, 3, 2, 1, 1e4, "-2", "lm"),
    ("p5_s2_M2_novec", 7, 17,  23, 29, 5, 3, 2, 1, 0.5, "2", "lm"),
    ("p3_s4_M0_gaps", 6,  11,  23, 31, 3, 2, 4, 1, 1e4, "abs", "lm"),
    ("T96_fallback",  98, 120, 15, 19, 7, 3, 4, 1, 0.0, "-2", "lm"),
    ("n2_1024",       6,  1026, 15, 15, 7, 3, 4, 1, 0.0, "-2", "lm"),
    ("st2_pt4",       14, 25,  19, 23, 5, 4, 2, 2, 1e4, "mse", "lm"),
    ("pt1_s1",        4,  6,   9,  11, 3, 1, 1, 1, 0.0, "-2", "lm"),
    ("direct_unfit",  9,  14,  24, 30, 7, 2, 4, 2, 0.3, "0", "direct"),
    ("p8_M2_even",    6,  15,  28, 32, 8, 2, 4, 1, 1e4, "-2", "lm"),
    # >= 40 query frames and a candidate set covered in <= 2 sweeps: the 4x8-tile strip kernel
    ("p11_tile8_tma", 44, 100, 27, 36, 11, 3, 4, 1, 0.0, "-2", "lm"),      # 9 groups (odd): TMA staging, 2 row groups
    ("p7_tile8",      42, 70,  23, 32, 7, 3, 4, 1, 1e4, "-2", "lm"),       # 6 groups: LDGSTS staging, 1 row group
    ("p8_tile8_rem0", 41, 300, 28, 36, 8, 2, 4, 1, 0.0, "0", "lm"),        # no padding lane, two sweeps, p == 2 s
    ("p4_tile8_tma",  43, 60,  16, 24, 4, 3, 4, 1, 1e4, "abs", "lm"),      # 3 groups: TMA staging, p == s
    ("p15_tile8_M3",  20, 45,  31, 40, 15, 3, 4, 1, 0.0, "-2", "lm"),      # 3 row groups in the local-memory ring, few frames
    # BASELINE config 5 corners: 4096 candidates, T = 96 with the larger patches, the other-view config at T = 48
    ("n2_4096",       6,  4098, 15, 15, 7, 3, 4, 1, 0.0, "-2", "lm"),
    ("n2_4096_p11",   5,  4098, 15, 19, 11, 3, 4, 1, 1e4, "-2", "lm"),
    ("T96_p11",       98, 130, 19, 23, 11, 3, 4, 1, 0.0, "-2", "lm"),
    ("T96_p15",       98, 110, 23, 27, 15, 3, 4, 1, 1e4, "-2", "lm"),
    ("T96_p3_s2",     98, 100, 11, 15, 3, 3, 2, 1, 1e4, "-2", "lm"),
    ("T48_p3_s2",     50, 66,  21, 27, 3, 3, 2, 1, 1e4, "-2", "lm"),
    # small-patch kernel (patchnn_diag.cuh: diagonal sums / arg-min in registers): every instantiation, with and without
    # the column-minimum normaliser, ragged query / candidate blocks, more than one candidate sweep, odd window starts
    ("diag_p3_s2_a0",   11, 37,  17, 25, 3, 3, 2, 1, 0.0, "-2", "lm"),     # M = 1, one extra row, alpha normaliser
    ("diag_p3_s3",      9,  20,  18, 24, 3, 3, 3, 1, 1e4, "-2", "lm"),     # M = 1, no extra row
    ("diag_p4_s4",      14, 23,  20, 28, 4, 3, 4, 1, 0.5, "0", "lm"),      # M = 1, full chunk
    ("diag_p4_s3",      7,  19,  22, 25, 4, 3, 3, 1, 1e4, "abs", "lm"),    # M = 1, one extra row, full chunk
    ("diag_p4_s2",      10, 41,  16, 22, 4, 3, 2, 1, 0.0, "-2", "lm"),     # M = 2 (ring in local memory)
    ("diag_p3_s1",      6,  13,  9,  12, 3, 3, 1, 1, 1e4, "-2", "lm"),     # M = 3
    ("diag_2sweeps",    6,  300, 11, 13, 3, 3, 2, 1, 0.0, "-2", "lm"),     # 298 candidates: two sweeps of 256
    ("diag_T96_2sw",    98, 150, 9,  11, 3, 3, 2, 1, 0.0, "-2", "lm"),     # 96 query positions: 24 x 16 threads, two sweeps
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_loss_kernels_match_oracle(case):
    import videoloop3d_b200 as V
    name, t, F_, h, w, p, pt, s, st, alpha, rou, cls = case
    dev = torch.device("cuda:0")
    # a literal, reproducible seed per case (Python's hash() of a str is randomised per process)
    g = torch.Generator().manual_seed(100 + [c[0] for c in CASES].index(name))
    y = torch.rand(1, 3, F_, h, w, generator=g)
    y = (y + y.roll(1, 2) + y.roll(1, 4)) / 3
    x = torch.rand(1, 3, t, h, w, generator=g) * 0.5 + 0.5 * y[:, :, torch.randint(0, F_, (t,), generator=g)]
    cfg = dict(patch_size=p, patcht_size=pt, stride=s, stridet=st, alpha=alpha, rou=rou, scaling=0.15)
    xo = x.double().requires_grad_(True)
    fn = LL.gpnn_lowmem if cls == "lm" else LL.gpnn_direct
    loss_o, aux = fn(xo, y.double(), macro_block=33, **cfg)
    (go,) = torch.autograd.grad(loss_o, xo)

    lossobj = V.Patch3DGPNNLowMemLoss() if cls == "lm" else V.Patch3DGPNNDirectLoss()
    # feed a non-contiguous (permuted) view, as MPMeshVid.forward does (MPV.py:506)
    xc = x.to(dev).permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4).requires_grad_(True)
    yc = y.to(dev).permute(0, 2, 1, 3, 4).contiguous().permute(0, 2, 1, 3, 4)
    kw = dict(macro_block=33) if cls == "lm" else {}
    loss_c = lossobj(xc, yc, **kw, **cfg)
    loss_c.backward()
    nn_c = lossobj.last_nn.cpu().long()
    assert nn_c.shape == aux["nn"].shape
    nn_over = None
    if not torch.equal(nn_c, aux["nn"]):
        xs, ys = aux["x"].detach(), y.double()[..., :aux["x"].shape[-2], :aux["x"].shape[-1]]
        mism, bad = LL.tie_margin_ok(xs, ys, aux["nn"], nn_c, p, pt, s, st, alpha)
        assert bad == 0 and mism <= 2, (mism, bad)
        loss_o, aux = fn(xo, y.double(), macro_block=33, nn_override=nn_c, **cfg)       # continue with CUDA's tie choice
        (go,) = torch.autograd.grad(loss_o, xo)
    assert abs(float(loss_c) - float(loss_o)) < 1e-4 * abs(float(loss_o)) + 1e-9
    assert float((lossobj.last_y2x.cpu().double() - aux["y2x"]).abs().max()) < 1e-5
    assert torch.equal(lossobj.last_weight.cpu(), aux["weight"].float())          # clamp_min(1e-10) is an fp32 value
    gc = xc.grad.cpu().double()
    assert float((gc - go).abs().max()) < 1e-4 * float(go.abs().max()) + 1e-12
