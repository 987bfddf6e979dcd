"""GPU tuning / A-B aid for the composite kernels (development tool, not part of the product path).

    python scripts/tune_composite.py [--workload step360p] [--frames 48] [--reps 5]

Times the pure render and the backward for a list of knob settings (VL3D_COMPOSITE_V1, VL3D_FWD_TF,
VL3D_BWD_TF) on the dense bench model and checks every variant against the first one
(rendered RGB, texel gradients, regulariser sums)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from videoloop3d_b200 import ops  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="step360p")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-smooth", action="store_true")
    ap.add_argument("--only", default="", help="comma list of variant names")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    wl = dict(bench.WORKLOADS[a.workload])
    T = a.frames or wl["T"]
    H, W = wl["H"], wl["W"]
    m = bench.build_model(wl, dev, T)
    atlas_dyn, atlas = m._texels()
    ext, intr = bench.view_for(wl)
    view = m.make_view(H, W, ext.reshape(4, 4).double().numpy(), intr)
    pack = m._pack
    g = torch.Generator(device=dev).manual_seed(5)
    rgb = torch.empty((T, 3, H, W), device=dev)
    grad_rgb = torch.randn((T, 3, H, W), device=dev, generator=g) * 1e-6
    w_smooth = None if a.no_smooth else torch.tensor([1e-9, 1.2e-9, 2e-9, 2.2e-9], device=dev)
    g_dyn = torch.empty_like(atlas_dyn.data)
    g_sta = torch.zeros_like(atlas.data)
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    fwd_b, bwd_b = bench.algorithmic_bytes(wl, T)

    variants = [
        ("v1", dict(VL3D_COMPOSITE_V1="1")),
        ("lean", dict()),
        ("lean notma", dict(VL3D_TMA="0", VL3D_TMA_BWD="0")),
        ("bwd notma", dict(VL3D_TMA_BWD="0")),
        ("tma s2", dict(VL3D_TMA_STAGES="2")),
        ("tma s4", dict(VL3D_TMA_STAGES="4")),
        ("tma tf2 s4", dict(VL3D_TMA_TF="2", VL3D_TMA_STAGES="4")),
        ("tma tf4 s2", dict(VL3D_TMA_TF="4", VL3D_TMA_STAGES="2")),
        ("tma tf4", dict(VL3D_TMA_TF="4")),
        ("lean fwd2", dict(VL3D_FWD_TF="2", VL3D_TMA="0")),
        ("lean fwd4", dict(VL3D_FWD_TF="4", VL3D_TMA="0")),
        ("lean bwd1", dict(VL3D_BWD_TF="1")),
        ("lean bwd3", dict(VL3D_BWD_TF="3")),
        ("lean bwd4", dict(VL3D_BWD_TF="4")),
        ("lean noRED", dict(VL3D_BWD_NORED="1")),
    ]
    if a.only:
        keep = set(a.only.split(","))
        variants = [v for v in variants if v[0] in keep]
    knobs = ("VL3D_COMPOSITE_V1", "VL3D_FWD_TF", "VL3D_FWD_MINB", "VL3D_BWD_TF", "VL3D_BWD_NORED", "VL3D_TMA", "VL3D_TMA_TF", "VL3D_TMA_STAGES", "VL3D_TMA_BWD")
    ref = None
    print(f"{a.workload}: {H}x{W}, D={wl['D']}, T={T}; algorithmic GB fwd {fwd_b / 1e9:.2f} bwd {bwd_b / 1e9:.2f}")
    for name, env in variants:
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update(env)

        def fwd():
            ops.composite_fwd(view, pack, atlas_dyn.data, atlas.data, None, T, 0, rgb_out=rgb)

        def bwd():
            ops.composite_bwd(view, pack, atlas_dyn.data, atlas.data, None, T, 0, grad_rgb, rgb, w_smooth, g_dyn, g_sta,
                              smooth_sums=None if a.no_smooth else sums)

        fmin, favg = timed(fwd, a.reps)
        bmin, bavg = timed(bwd, a.reps)
        g_dyn.zero_(); g_sta.zero_(); sums.zero_()
        fwd(); bwd()
        torch.cuda.synchronize()
        cur = (rgb.clone(), g_dyn[: min(T, 2)].clone(), sums.clone())
        msg = ""
        if ref is None:
            ref = cur
        elif "noRED" not in name:
            e_rgb = float((cur[0] - ref[0]).abs().max() / ref[0].abs().max())
            e_g = float((cur[1] - ref[1]).abs().max() / ref[1].abs().max())
            e_s = float(((cur[2] - ref[2]).abs() / ref[2].abs().clamp_min(1e-30)).max())
            msg = f"  vs {variants[0][0]}: rgb {e_rgb:.1e} grad {e_g:.1e} sums {e_s:.1e}"
        print(f"{name:12s} fwd {fmin:7.3f} ms ({fwd_b / fmin / 1e6:6.0f} GB/s)  bwd {bmin:7.3f} ms ({bwd_b / bmin / 1e6:6.0f} GB/s)"
              f"  [avg {favg:.3f} / {bavg:.3f}]{msg}", flush=True)


if __name__ == "__main__":
    main()
