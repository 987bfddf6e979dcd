"""TEST INFRASTRUCTURE — times the UNMODIFIED reference (behind oracle/ref_shims) next to the oracle port on the same
sample of the bench workload, on this container's host cores (needs /root/reference; cannot run on the GPU box).

    python oracle/time_reference_vs_port.py [--out profiles/r02_cpu_reference_vs_port.json]

Sample: one 90x160 patch (1/64 of the 720p frame), D=32, T=48, F=66 target frames, reference-view loss config, one
optimisation step = forward (render + gpnn_lm + smoothness) + backward + Adam.  The vertex mesh is 9x16 for both arms
(the naive rasteriser stand-in for pytorch3d is O(faces x pixels): its time is reported separately and is not part
of either arm's step time)."""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import mpv_oracle as MO  # noqa: E402
from oracle import ref_env  # noqa: E402
from oracle.make_golden import _batched, _load_state_into_reference, _view  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_cpu_reference_vs_port.json"))
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    ref_env.enable()
    import MPV  # the reference, unmodified
    import utils as ref_utils
    H, W, D, T, F, hv, wv = 90, 160, 32, 48, 66, 9, 16
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    args = ref_env.make_args(mpi_d=D, mpi_h_verts=hv, mpi_w_verts=wv, atlas_grid_h=4, mpv_frm_num=T, mpi_h_scale=1.0,
                             mpi_w_scale=1.0, add_intrin_noise=False)
    f = 0.8 * W
    ref_intrin = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(2)
    m = MPV.MPMeshVid(args, H, W, np.eye(4, dtype=np.float32), ref_intrin, 1.0, 10.0)
    st = MO.dense_state(H, W, D, hv, wv, 4, T, 1.0, 10.0, 1.0, 1.0, seed=2)
    st.atlas = st.atlas[:, :, :1, :1].clone()
    _load_state_into_reference(m, st)
    ext, intr = _view(2, H, W)
    g = torch.Generator().manual_seed(3)
    res = torch.rand(1, F, 3, H, W, generator=g)
    cfg = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=11, patcht_size=3, stride=4, stridet=1, alpha=0.0, rou="-2",
               scaling=0.1, dist_fn="mse", macro_block=65, factor=1)

    # time the rasteriser stand-in separately by wrapping the reference's own call site (utils.py:31-70)
    raster_s = [0.0]
    orig_forward = ref_utils.SimpleRasterizer.forward

    def timed_forward(self, *x, **k):
        t0 = time.perf_counter()
        r = orig_forward(self, *x, **k)
        raster_s[0] += time.perf_counter() - t0
        return r

    ref_utils.SimpleRasterizer.forward = timed_forward
    m.train()
    opt = torch.optim.Adam([m.atlas, m.atlas_dyn], lr=0.005, betas=(0.9, 0.999), eps=6e-8)
    ref_times, ref_raster = [], []
    for i in range(1 + a.reps):
        raster_s[0] = 0.0
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, extra = m(H, W, ext, intr, res=res, losscfg=_batched(cfg))
        loss = extra.pop("swd").mean()
        for k, v in extra.items():
            w_ = getattr(args, f"{k}_loss_weight")
            if w_ > 0:
                loss = loss + v.mean() * w_
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i > 0:
            ref_times.append(dt - raster_s[0])
            ref_raster.append(raster_s[0])
        if i == 0:
            ref_loss = float(loss)
    ref_utils.SimpleRasterizer.forward = orig_forward

    # the oracle port, same state / view / target (bench.py's cpu_reference_step does exactly this)
    ad = st.atlas_dyn.float().requires_grad_(True)
    at = st.atlas.float().requires_grad_(True)
    mom, var = torch.zeros_like(ad), torch.zeros_like(ad)
    port_times = []
    for i in range(1 + a.reps):
        t0 = time.perf_counter()
        extra, _ = MO.forward_train(st, H, W, ext, intr, res, cfg, dtype=torch.float32, atlas=at, atlas_dyn=ad, nn_mode="ref32")
        loss = MO.total_loss(extra)
        ad.grad = None
        loss.backward()
        with torch.no_grad():
            p, mom, var = MO.adam_step(ad.detach(), ad.grad, mom, var, i + 1, 0.005)
            if i == 0:
                port_loss = float(loss)
            ad.data.copy_(p)
        if i > 0:
            port_times.append(time.perf_counter() - t0)
    out = {"sample": f"{H}x{W} patch, D={D}, T={T}, F={F}, mesh {hv}x{wv}, ref-view loss cfg, fwd + bwd + Adam, fp32",
           "cores": threads, "reference_verbatim_s": float(np.mean(ref_times)), "reference_raster_shim_s": float(np.mean(ref_raster)),
           "oracle_port_s": float(np.mean(port_times)), "first_step_loss": {"reference": ref_loss, "port": port_loss}}
    out["ratio_port_over_reference"] = out["oracle_port_s"] / out["reference_verbatim_s"]
    with open(a.out, "w") as fo:
        json.dump(out, fo, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
