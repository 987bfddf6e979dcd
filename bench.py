#!/usr/bin/env python
"""bench.py — VideoLoop3D stage-2 hot path on B200: render + looping loss + backward + Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Prints ONE JSON line (rank 0).  Metric (BASELINE.json): "MPV render+loop-loss optim steps/sec
@ D=32,T=48,720p".  Workload `step720p` = BASELINE configs[2] (the configuration the metric is quoted on;
it contains configs[1], the render fwd+bwd, as a sub-part that is also reported):
dense synthetic MPV, D=32 planes, T=48 frames, 720x1280, 36x64 vertex mesh, 1 texel : 1 pixel, reference-view
loss config of configs/mpv_base.txt (p=11, pt=3, s=4, alpha=0, rou=-2, gain 3.5), rgb/a smoothness 0.2,
scale-invariant gain, F=258 target frames (n2=256), Adam(eps=6e-8) over every texel.
With --gpus N the T frames are sharded over N ranks (strong scaling: the step is the same).

`--impl reference`: the reference's own algorithm on the host CPU (oracle port of its PyTorch path; the
reference itself needs pytorch3d/unfoldNd which cannot be installed here), bounded sample, same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "MPV render+loop-loss optim steps/sec @ D=32,T=48,720p"
WORKLOADS = {
    # name: H, W, D, T, F, hv, wv
    "step720p": dict(H=720, W=1280, D=32, T=48, F=258, hv=36, wv=64),
    "step360p": dict(H=360, W=640, D=32, T=48, F=258, hv=36, wv=64),
    "patch180": dict(H=180, W=320, D=32, T=48, F=258, hv=36, wv=64),      # the reference's own step shape
    # SURVEY §8(d) "sparse": Bernoulli(0.25) tile occupancy, half static / half dynamic, 32x32-texel tiles
    "sparse720p": dict(H=720, W=1280, D=32, T=48, F=258, hv=36, wv=64, sparse=dict(tile=32, occupancy=0.25, dyn_frac=0.5)),
    "tiny": dict(H=45, W=80, D=8, T=6, F=12, hv=6, wv=9),
}
CPU_SAMPLE = dict(H=90, W=160, D=32, T=48, F=258, hv=36, wv=64)             # 1/64 of the 720p frame


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def view_for(wl, jitter=(0.21, -0.13)):
    """Target camera of SURVEY §8(d): 0.05 rad about y + translation (0.08,-0.03,0.02)*near, jittered
    principal point."""
    H, W = wl["H"], wl["W"]
    c, s = np.cos(0.05), np.sin(0.05)
    ext = np.eye(4, dtype=np.float32)
    ext[:3, :3] = [[c, 0, s], [0, 1, 0], [-s, 0, c]]
    ext[:3, 3] = [0.08, -0.03, 0.02]
    f = 0.8 * W
    intr = np.array([[f, 0, W / 2 + jitter[0]], [0, f, H / 2 + jitter[1]], [0, 0, 1]], dtype=np.float32)
    return torch.from_numpy(ext)[None], torch.from_numpy(intr)[None]


def make_args(wl, **kw):
    from videoloop3d_b200 import default_args
    return default_args(mpi_d=wl["D"], mpi_h_verts=wl["hv"], mpi_w_verts=wl["wv"], atlas_grid_h=4 if wl["D"] % 4 == 0 else 1,
                        mpi_h_scale=1.0, mpi_w_scale=1.0, mpv_frm_num=1, **kw)


def build_model(wl, device, frames, seed=2, first_frame=0):
    """Dense model as MPMeshVid.__init__ lays it out; the (frames,4,Hd,Wd) dynamic atlas is generated on the
    device (N(0,1) rgb logits, N(-1,1) alpha logits) straight into the RGBA-interleaved layout.  Frame t of the
    video is seeded by (seed, t) alone, so a rank that holds frames [first_frame, first_frame + frames) builds
    exactly the texels the single-GPU model has there (the loss is then comparable across N)."""
    from videoloop3d_b200 import MPMeshVid
    H, W = wl["H"], wl["W"]
    args = make_args(wl)
    f = 0.8 * W
    ref_intrin = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(seed)
    m = MPMeshVid(args, H, W, np.eye(4, dtype=np.float32), ref_intrin, 1.0, 10.0)
    hd, wd = m.atlas_dyn.shape[-2:]
    m.atlas.data = m.atlas.data[:, :, :1, :1].clone()             # dummy static atlas (MPV.py:266)
    m.atlas_dyn.data = m.atlas_dyn.data[:, :, :1, :1].clone()
    m = m.to(device)
    if wl.get("sparse"):
        from videoloop3d_b200.testing import cull_to_tiles
        sp = wl["sparse"]
        return cull_to_tiles(m, sp["tile"], sp["occupancy"], sp["dyn_frac"], frames, seed=seed, device=device,
                             first_frame=first_frame)
    g = torch.Generator(device=device)
    tex = torch.empty((frames, hd, wd, 4), dtype=torch.float32, device=device)
    for t in range(frames):
        g.manual_seed(seed * 100003 + first_frame + t)
        tex[t].normal_(generator=g)
    tex[..., 3] -= 1.0
    m.atlas_dyn.data = tex.permute(0, 3, 1, 2)                     # logical (T,4,Hd,Wd), channels_last memory
    m.frm_num = frames
    m.invalidate_geometry()
    return m


def make_target(wl, device=None, seed=3, pinned=False, as_bytes=False):
    """Target video: U(0,1) noise low-pass filtered in time (box 9), quantised to 8 bits like a decoded video
    (MVVidPatchDataset holds `uint8 frames / 255`, train_3dvid.py:52-54).  (1,F,3,H,W) float32 = bytes / 255, or the
    bytes themselves with `as_bytes`."""
    F_, H, W = wl["F"], wl["H"], wl["W"]
    dev = device if (device is not None and device.type == "cuda") else None
    g = torch.Generator(device=dev).manual_seed(seed) if dev is not None else torch.Generator().manual_seed(seed)
    raw = torch.rand((F_ + 8, 3, H, W), generator=g, device=dev)
    cs = torch.cumsum(raw, 0)
    del raw
    u8 = torch.empty((F_, 3, H, W), dtype=torch.uint8, device=dev)
    u8[0] = (cs[8] / 9 * 255).round().clamp_(0, 255).to(torch.uint8)
    for a in range(1, F_, 32):                                   # in slabs: no second full-size float temporary
        b = min(a + 32, F_)
        u8[a:b] = ((cs[a + 8:b + 8] - cs[a - 1:b - 1]) / 9 * 255).round().clamp_(0, 255).to(torch.uint8)
    del cs
    if as_bytes:
        out = u8[None]
        return out.pin_memory() if (pinned and dev is None) else out
    res = torch.empty((1, F_, 3, H, W), dtype=torch.float32, device=dev)
    for a in range(0, F_, 32):
        res[0, a:a + 32] = u8[a:a + 32].float() / 255
    return res.pin_memory() if (pinned and dev is None) else res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the committed ncu capture of this
# same workload (source file named in the line); None where nothing was captured.  A bench run cannot read DRAM
# counters itself, so this is the capture's value, not a measurement of the run that prints it.
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the named ncu capture of this very
# command (a capture cannot run inside the timed bench); keyed by what the run actually used.
NCU_TRAFFIC_BYTES = {
    ("step720p", 1, "fused_bwd_adam"): (117.592024e9 + 94.954704e9,
                                        "profiles/r03p_fused_720p_metrics.csv (generic schedule, compressible gradient buffer)"),
    ("step720p", 1, "fused_bwd_adam", "plain"): (139.307588e9 + 112.821573e9,
                                                 "profiles/r03h_fused_720p_compress0.csv (generic schedule, plain gradient buffer)"),
    ("step720p", 1, "composite_bwd"): (42.817052e9 + 20.455665e9, "profiles/r01c_ncu_full_step720p.md"),
    ("step720p", 1, "composite_fwd"): (21.071395e9 + 0.553144e9, "profiles/r01c_ncu_full_step720p.md"),
}


def algorithmic_bytes(wl, frames):
    """each live texel touched once per pass, geometry recomputed in-kernel.
    Dense 1:1 case N_tex = D*H*W per frame: fwd 16 B/texel + 12 B/pixel out; bwd re-reads the atlas and
    writes the gradient (32 B/texel) + reads dL/drgb (12 B/pixel)."""
    px = wl["H"] * wl["W"]
    ntex = wl["D"] * px
    fwd = frames * (16 * ntex + 12 * px)
    bwd = frames * (32 * ntex + 12 * px)
    return fwd, bwd


def fused_algorithmic_bytes(wl, frames):
    """Fused backward + Adam: every texel's p, m, v read once and written once (the backward and Adam share the read
    of p; the gradient never has to leave the chip) + dL/drgb and rgb read per pixel."""
    px = wl["H"] * wl["W"]
    return frames * (96 * wl["D"] * px + 24 * px)


# --------------------------------------------------------------------------------------------------
def cpu_reference_step(sample, threads, steps, warmup=1):
    """The reference's algorithm (oracle port of its PyTorch CPU path, fp32) on a bounded sample:
    forward (render + gpnn_lm + smoothness) + backward + Adam.  Returns seconds per sample step."""
    from oracle import mpv_oracle as MO
    torch.set_num_threads(threads)
    st = MO.dense_state(sample["H"], sample["W"], sample["D"], sample["hv"], sample["wv"], 4 if sample["D"] % 4 == 0 else 1,
                        sample["T"], 1.0, 10.0, 1.0, 1.0, seed=2)
    st.atlas = st.atlas[:, :, :1, :1].clone()
    ext, intr = view_for(sample)
    res = make_target(sample)
    cfg = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=11, patcht_size=3, stride=4, stridet=1, alpha=0.0,
               rou="-2", scaling=0.1, dist_fn="mse", macro_block=65, factor=1)
    geo = MO.geometry(st, sample["H"], sample["W"], ext, intr)
    ad = st.atlas_dyn.float().requires_grad_(True)
    a = st.atlas.float().requires_grad_(True)
    mom, var = torch.zeros_like(ad), torch.zeros_like(ad)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        MO_geo = MO.geometry(st, sample["H"], sample["W"], ext, intr)          # the reference rasterises every step
        extra, _ = MO.forward_train(st, sample["H"], sample["W"], ext, intr, res, cfg, dtype=torch.float32, atlas=a,
                                    atlas_dyn=ad, nn_mode="ref32")
        loss = MO.total_loss(extra)
        ad.grad = None
        loss.backward()
        with torch.no_grad():
            p, mom, var = MO.adam_step(ad.detach(), ad.grad, mom, var, i + 1, 0.01)
            ad.data.copy_(p)
        times.append(time.perf_counter() - t0)
        del MO_geo
    return float(np.mean(times[warmup:]))


def gpu_torch_reference_step(wl, dev, steps):
    """The reference's torch operator sequence (oracle/torch_ref_ops.py) on the GPU, SURVEY §8(d) "reference GPU
    baseline": the frame is covered by non-overlapping patches of the reference's own step shape (180x320,
    configs/mpv_base.txt:21-24), gradients accumulated, one Adam step = one full-frame-equivalent step.
    The rasteriser stand-in runs on the host beforehand and is NOT timed.  Adam is timed after the patches (its
    state is allocated only then, so that the patches' autograd temporaries and the optimiser state never
    coexist: 5 x 22.6 GB during a patch, 6 x 22.6 GB during Adam at 720p).  Returns (ms patches, ms adam, n)."""
    from oracle import mpv_oracle as MO
    from oracle import torch_ref_ops as RO
    H, W, D, T = wl["H"], wl["W"], wl["D"], wl["T"]
    ph, pw = min(H, 180), min(W, 320)
    st = MO.dense_state(H, W, D, wl["hv"], wl["wv"], 4 if D % 4 == 0 else 1, 1, 1.0, 10.0, 1.0, 1.0, seed=2)
    hd, wd = st.atlas_dyn.shape[-2:]
    st.atlas = st.atlas[:, :, :1, :1].clone()
    ext, intr = view_for(wl)
    tiles = []
    for hs in range(0, H - ph + 1, ph):
        for ws in range(0, W - pw + 1, pw):
            k = intr.clone()
            k[0, 0, 2] -= ws                                        # get_new_intrin, utils.py:196-200
            k[0, 1, 2] -= hs
            tiles.append((hs, ws, RO.raster_tables(st, ph, pw, ext, k, dev)))
    g = torch.Generator(device=dev).manual_seed(2)
    atlas_dyn = torch.empty((T, 4, hd, wd), dtype=torch.float32, device=dev)
    for t in range(T):
        atlas_dyn[t].normal_(generator=g)
    atlas_dyn[:, 3] -= 1.0
    atlas_dyn.requires_grad_(True)
    atlas = torch.zeros((1, 4, 1, 1), device=dev, requires_grad=True)
    res = make_target(wl, dev)
    cfg = dict(loss_name="gpnn_lm", loss_gain=3.5, patch_size=11, patcht_size=3, stride=4, stridet=1, alpha=0.0,
               rou="-2", scaling=0.1, dist_fn="mse", macro_block=65, factor=1)
    opt = torch.optim.Adam([atlas_dyn], lr=1e-4, betas=(0.9, 0.999), eps=6e-8, foreach=False)

    def patches(which):
        for hs, ws, tabs in which:
            crop = res[..., hs:hs + ph, ws:ws + pw]
            total, _, _ = RO.forward_train_ops(atlas, atlas_dyn, tabs, ph, pw, crop, cfg, D)
            total.backward()

    patches(tiles[:1])                                              # warm-up (allocator, cuBLAS handles)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    for _ in range(steps):
        patches(tiles)
    e[1].record()
    torch.cuda.synchronize()
    opt.step()                                                      # allocates exp_avg / exp_avg_sq
    torch.cuda.synchronize()
    e[2].record()
    for _ in range(steps):
        opt.step()
    e[3].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) / steps, e[2].elapsed_time(e[3]) / steps, len(tiles), (ph, pw)


def run_reference_cuda(args):
    """`--impl reference --ref-device cuda`: not the driver's reference arm (that one is the host-CPU run below),
    but the comparison BASELINE.json's north_star names: the reference's PyTorch step on the same single B200."""
    wl = WORKLOADS[args.workload]
    base = {"impl": "reference", "metric": METRIC, "unit": "steps/s", "n_gpus": 1, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "device": "cuda:0 (torch eager, allow_tf32=False)"}}
    if not torch.cuda.is_available():
        emit_line({**base, "unavailable": "--ref-device cuda needs a GPU"})
        return
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    steps = max(1, min(args.steps, 2))
    try:
        ms_p, ms_a, n, (ph, pw) = gpu_torch_reference_step(wl, torch.device("cuda", 0), steps)
    except torch.cuda.OutOfMemoryError as ex:
        emit_line({**base, "unavailable": f"out of memory: {str(ex)[:120]}"})
        return
    ms = ms_p + ms_a
    emit_line({**base, "value": 1000.0 / ms, "steps": steps, "warmup": 1, "ms_per_step": ms,
                      "parts_ms": {"patches_fwd_bwd": ms_p, "adam": ms_a},
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                      "sample": f"{n} patches of {ph}x{pw} (the reference's step shape) covering {wl['H']}x{wl['W']}, "
                                f"gradients accumulated, one torch.optim.Adam step; reference operator sequence "
                                f"(grid_sample / masked_scatter / cumprod / unfold / bmm / index_add), host "
                                f"rasteriser stand-in not timed",
                      "gpu_launches": 0})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_cuda(args)
    threads = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    sample = dict(CPU_SAMPLE) if args.workload == "step720p" else dict(wl)
    ratio = (wl["H"] * wl["W"]) / (sample["H"] * sample["W"])
    sec = cpu_reference_step(sample, threads, max(1, args.steps), warmup=min(1, args.warmup))
    value = 1.0 / (sec * ratio)
    desc = (f"{sample['H']}x{sample['W']} patch (1/{ratio:.0f} of the frame), D={sample['D']}, T={sample['T']}, "
            f"F={sample['F']}; {sec:.2f} s per sample step, scaled x{ratio:.0f} to the full frame")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sec * ratio, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)



# --------------------------------------------------------------------------------------------------
# extra measurements carried by the default line
# --------------------------------------------------------------------------------------------------
def time_steps(step, call, steps, warmup, barrier):
    """ms per step (CUDA events on the launching stream, max over ranks) + per-kernel ms of `steps` calls."""
    import torch.distributed as dist
    for _ in range(warmup):
        call()
    barrier()
    n0 = {k: len(v) for k, v in (step.timers or {}).items()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = call()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    kms = {}
    for k, evs in (step.timers or {}).items():
        evs = evs[n0.get(k, 0):]
        if evs:
            kms[k] = round(sum(a.elapsed_time(b) for a, b in evs) / len(evs), 4)
    return float(ms) / steps, kms, out


def loss_sweep(dev, extents=((180, 320),), ps=(7, 11, 15), Ts=(24, 48, 96), n2s=(256, 1024, 4096), reps=2):
    """BASELINE config 5: the looping-loss kernels (NN search + vote / robust loss / gradient) over patch size, video
    length and candidate-set size; reference-view config otherwise (pt=3, s=4, st=1, alpha=0, rou=-2).  ms per call."""
    from videoloop3d_b200 import ops
    out = {}
    g = torch.Generator(device=dev).manual_seed(5)
    for (h, w) in extents:
        y_all = torch.rand((max(n2s) + 2, 3, h, w), device=dev, generator=g)
        for T in Ts:
            x = torch.rand((T + 2, 3, h, w), device=dev, generator=g)
            xs = torch.ones(1, device=dev)
            for n2 in n2s:
                y = y_all[:n2 + 2]
                for p in ps:
                    desc = ops.make_loss_desc(x.shape, (x.stride(0), x.stride(1), x.stride(2)), y.shape,
                                              (y.stride(0), y.stride(1), y.stride(2)), p, 3, 4, 1, 0.0)
                    nn = torch.empty((desc.ho, desc.wo, desc.n1), dtype=torch.int32, device=dev)
                    grad = torch.empty_like(x)
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    for r in range(1 + reps):
                        if r == 1:
                            ev[0].record()
                        ops.patchnn_search(desc, x, None, y, nn_out=nn)
                    ev[1].record()
                    for r in range(reps):
                        ops.vote_loss(desc, x, xs, y, nn, "-2", 0.1, 1.0, (T + 2, h, w), grad_out=grad)
                    ev[2].record()
                    torch.cuda.synchronize()
                    out[f"{h}x{w}_p{p}_T{T}_n{n2}"] = {"search_ms": round(ev[0].elapsed_time(ev[1]) / reps, 3),
                                                       "vote_ms": round(ev[1].elapsed_time(ev[2]) / reps, 3)}
        del y_all
    return out


def config0_static_render(dev):
    """BASELINE configs[0]: single-view MPI render, D=8 planes, 256x256, static frame — the reference's CPU path
    (oracle port, all host threads) next to the same render through libvl3d on the GPU."""
    from oracle import mpv_oracle as MO
    from videoloop3d_b200.testing import model_from_tensors
    H = W = 256
    st = MO.sparse_state(H, W, 8, 9, 9, 1, 1.0, 10.0, tile=32, occupancy=1.0, dyn_frac=0.0, h_scale=1.0, w_scale=1.0, seed=2)
    wl = dict(H=H, W=W)
    ext, intr = view_for(wl)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        rgb_o, _ = MO.render(st, H, W, ext, intr, [0], dtype=torch.float32)
    cpu_ms = (time.perf_counter() - t0) / n * 1e3
    tens = dict(verts=st.verts, planedepth=st.planedepth, faces=st.faces, faces_dyn=st.faces_dyn, uvs=st.uvs,
                uvs_dyn=st.uvs_dyn, uvfaces=st.uvfaces, uvfaces_dyn=st.uvfaces_dyn, atlas=st.atlas, atlas_dyn=st.atlas_dyn,
                ref_extrin=st.ref_extrin, ref_intrin=st.ref_intrin, mpi_d=8, hv=9, wv=9)
    m = model_from_tensors(tens, H, W, dev)
    m.eval()
    with torch.no_grad():
        for _ in range(3):
            rgb, _ = m(H, W, ext, intr, ts=[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            rgb, _ = m(H, W, ext, intr, ts=[0])
        e1.record()
        torch.cuda.synchronize()
    err = float((rgb[0].permute(1, 2, 0).cpu() - rgb_o[0]).abs().max())
    return {"workload": "D=8, 256x256, static atlas, one frame (BASELINE configs[0])", "cpu_ms": round(cpu_ms, 3),
            "cpu_kind": "port", "cores": os.cpu_count() or 1, "gpu_ms_per_call": round(e0.elapsed_time(e1) / 20, 4),
            "max_abs_diff": err}


def sharded_check(dev, group, world, rank):
    """Before timing at N > 1: two optimisation steps of a small dense model (the reference's step shape, 180x320,
    D=32, T=16) — sharded over the ranks vs. every rank computing the whole thing alone — must give the same frames,
    the same NN map and the same parameters on the frames a rank owns."""
    import torch.distributed as dist
    from videoloop3d_b200 import FusedLoopStep
    from videoloop3d_b200.train_step import loss_config
    wl = dict(H=180, W=320, D=32, T=16, F=34, hv=9, wv=16)
    T = wl["T"]
    t0, t1 = (T * rank) // world, (T * (rank + 1)) // world
    full = build_model(wl, dev, T, seed=5)
    shard = build_model(wl, dev, t1 - t0, seed=5, first_frame=t0)
    assert torch.equal(full.atlas_dyn.data[t0:t1], shard.atlas_dyn.data)
    cfg = loss_config(full.args, ref_view=True)
    ext, intr = view_for(wl)
    res = make_target(wl, dev)
    a = FusedLoopStep(full)
    b = FusedLoopStep(shard, group=group, global_frames=T)
    for _ in range(2):
        oa = a.step(wl["H"], wl["W"], ext, intr, res, cfg, 0.005)
        ob = b.step(wl["H"], wl["W"], ext, intr, res, cfg, 0.005)
    torch.cuda.synchronize()
    d = (full.atlas_dyn.data[t0:t1] - shard.atlas_dyn.data).abs()
    stats = torch.tensor([float(d.max()), float((d > 1e-5).float().mean()),
                          abs(float(oa["loss"]) - float(ob["loss"])) / abs(float(oa["loss"])),
                          float((a._buf["nn"] != b._buf["nn"]).sum())], device=dev, dtype=torch.float64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX, group=group)
    return {"workload": "dense 180x320, D=32, T=16, F=34, 2 steps", "max_abs_dparam": float(stats[0]),
            "frac_dparam_gt_1e-5": float(stats[1]), "loss_rel_diff": float(stats[2]), "nn_mismatches": int(stats[3])}


# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from videoloop3d_b200 import FusedLoopStep, _lib
    from videoloop3d_b200.train_step import loss_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the vl3d hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    extras = not args.quick and args.workload == "step720p"
    check = sharded_check(dev, group, world, rank) if (world > 1 and not args.quick) else None
    torch.cuda.empty_cache()
    line = measure_main(args, world, rank, local, dev, group, barrier, extras)
    torch.cuda.empty_cache()
    if extras:
        more = measure_extras(args, world, rank, dev, group, barrier)
        if rank == 0:
            line.update(more)
    if rank == 0:
        if check is not None:
            line["sharded_check"] = check
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_entry(args)
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_entry(args):
    wl = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample = dict(CPU_SAMPLE) if args.workload in ("step720p", "sparse720p") else dict(wl)
    ratio = (wl["H"] * wl["W"]) / (sample["H"] * sample["W"])
    sec = cpu_reference_step(sample, threads, 1, warmup=0)
    return {"value": 1.0 / (sec * ratio), "unit": "steps/s", "cores": threads, "kind": "port",
            "sample": f"oracle port of the reference's PyTorch CPU path on one {sample['H']}x{sample['W']} "
                      f"patch (1/{ratio:.0f} of the frame), D={sample['D']}, T={sample['T']}, F={sample['F']}: "
                      f"{sec:.2f} s, scaled x{ratio:.0f}"}


def measure_main(args, world, rank, local, dev, group, barrier, extras):
    """The headline workload: resident-input arm (`value`), end-to-end arm (`e2e`), per-kernel times, roofline."""
    import torch.distributed as dist
    from videoloop3d_b200 import FusedLoopStep, _lib
    from videoloop3d_b200.train_step import loss_config
    wl = WORKLOADS[args.workload]
    T = wl["T"]
    t0, t1 = (T * rank) // world, (T * (rank + 1)) // world
    model = build_model(wl, dev, t1 - t0, seed=2, first_frame=t0)
    margs = model.args
    if args.no_smooth:
        margs.rgb_smooth_loss_weight = margs.a_smooth_loss_weight = 0.0
    cfg = loss_config(margs, ref_view=True)
    ext, intr = view_for(wl)
    H, W = wl["H"], wl["W"]
    res_dev = make_target(wl, dev)
    lr = margs.lrate * 0.01
    step = FusedLoopStep(model, group=group, global_frames=T, timers=True, fused=args.fused, exchange=args.exchange,
                         gather_nn=False)
    ya, yb = step.band_rows(H, cfg)                                    # this rank's rows of the target (N > 1: its loss band)
    res_main = res_dev if (ya, yb) == (0, H) else res_dev[:, :, :, ya:yb].contiguous()

    # ---------------- resident-input arm (value) ----------------
    for _ in range(args.warmup):
        step.step(H, W, ext, intr, res_main, cfg, lr)
    barrier()
    n_warm_events = len(step.timers.get("composite_fwd", []))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step.step(H, W, ext, intr, res_main, cfg, lr)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.LAUNCHES - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    kernel_ms = step.timer_ms(skip=n_warm_events)
    final_loss = float(out["loss"])

    # ---------------- end-to-end arm: host buffers, H2D + D2H inside the timed region ----------------
    # The public call (FusedLoopStep.step, the run_iter equivalent) is fed from pinned host memory the way
    # train_3dvid.run_iter feeds it (`datainfo_.to(device)`, train_3dvid.py:215); the next item's copy is
    # issued on a side stream while the current step computes (double buffer), the loss is read back.
    # The host holds the video the way a decoder delivers it and MVVidPatchDataset(storage="uint8") keeps it: bytes;
    # `/ 255` happens on the device inside the step (vl3d_u8_to_unit: the same bits as the host conversion).
    # With N ranks the loss is sharded by pixel-row bands, so every rank's loader holds — and copies over its own PCIe
    # link — only the rows of the target video its band needs (FusedLoopStep.band_rows); h2d_bytes_per_step is the
    # total over all ranks.
    res_full_host = make_target(wl, None, seed=3, as_bytes=True)
    res_host = res_full_host[:, :, :, ya:yb].contiguous().pin_memory()    # this rank's share of the item
    del res_full_host
    bufs_u8 = [torch.empty(tuple(res_host.shape), dtype=torch.uint8, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ext_h, intr_h = ext.pin_memory(), intr.pin_memory()
    rows_all = torch.tensor([yb - ya], device=dev)
    if world > 1:
        dist.all_reduce(rows_all)
    h2d = wl["F"] * 3 * int(rows_all) * W + 1216 * world   # target bytes (all ranks together) + the view descriptors
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def load(i):
        """H2D of this rank's rows of the target video, as bytes, on the copy stream."""
        bufs_u8[i].copy_(res_host, non_blocking=True)

    def e2e_loop(n):
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.stream(copy_stream):
            load(0)
            ready[0].record()
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(done[nxt])             # buffer `nxt` was consumed by step i-1
                    load(nxt)
                    ready[nxt].record()
            # pose / intrinsics stay on the host: the view descriptor (plane homographies) is built there;
            # the step waits for the copy only where it first reads the target video (after the render)
            o = step.step(H, W, ext_h, intr_h, bufs_u8[cur], cfg, lr, res_ready=ready[cur])
            done[cur].record()
            loss_host.copy_(o["loss"].reshape(1), non_blocking=True)
        torch.cuda.synchronize()

    e2e_loop(min(2, args.warmup))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1000
    ems = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_ms = float(ems) / args.steps

    if rank == 0:
        peak, peak_src = measured_peaks()
        fwd_b, bwd_b = algorithmic_bytes(wl, t1 - t0)
        fused_b = fused_algorithmic_bytes(wl, t1 - t0)
        cand = {"composite_fwd": fwd_b, "composite_bwd": bwd_b, "fused_bwd_adam": fused_b}
        dom = max((k for k in cand if k in kernel_ms), key=lambda k: kernel_ms[k])
        dom_bytes = cand[dom]
        ach = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9
        fwd_ach = fwd_b / (kernel_ms["composite_fwd"] * 1e-3) / 1e9
        comp = {"fwd_GBps": fwd_ach, "fwd_frac": fwd_ach / peak}
        if "composite_bwd" in kernel_ms and "fused_bwd_adam" not in kernel_ms:
            bwd_ach = bwd_b / (kernel_ms["composite_bwd"] * 1e-3) / 1e9
            comp.update({"bwd_GBps": bwd_ach, "bwd_frac": bwd_ach / peak,
                         "render_fwd_bwd_steps_per_s": 1000.0 / (kernel_ms["composite_fwd"] + kernel_ms["composite_bwd"]
                                                                 + kernel_ms.get("grad_zero", 0.0))})
        if "fused_bwd_adam" in kernel_ms:
            f_ach = fused_b / (kernel_ms["fused_bwd_adam"] * 1e-3) / 1e9
            comp.update({"fused_bwd_adam_GBps": f_ach, "fused_bwd_adam_frac": f_ach / peak,
                         "fused_schedule": step.last_schedule.kind if step.last_schedule is not None else None})
        traffic_key = (args.workload, world, dom)
        if dom == "fused_bwd_adam" and not step.grad_compressed:
            traffic_key += ("plain",)
        comp["grad_buffer"] = "compressible device memory" if step.grad_compressed else "plain device memory"
        line = {
            "metric": METRIC, "value": 1000.0 / ms_per_step, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "H": H, "W": W, "planes": wl["D"], "frames": T,
                       "target_frames": wl["F"], "mesh": [wl["hv"], wl["wv"]], "atlas": "dense 1 texel:1 pixel",
                       "loss": "gpnn_lm p=11 pt=3 s=4 alpha=0 rou=-2 gain=3.5 + rgb/a smooth 0.2 + scale-invariant",
                       "optimizer": "Adam eps=6e-8 over all texels",
                       "parallelism": f"T-shard x{world} (render / backward / Adam by frames, looping loss by pixel-row bands)",
                       "exchange": (None if world == 1 else
                                    ("peer-memory stores (vl3d_copy_boxes into symmetric memory)" if step._peer is not None
                                     else "NCCL all-to-all")),
                       "l2": "inputs (>= 22 GB of texels per step) far exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": 1000.0 / e2e_ms, "unit": "steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms, "wall_ms_per_step": wall_ms / args.steps,
                    "api": "FusedLoopStep.step fed from pinned host memory holding the uint8 target video (double-buffered "
                           "H2D on a copy stream, /255 (vl3d_u8_to_unit) inside the step; with N ranks each rank copies only the "
                           "pixel rows of its loss band), loss read back"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": NCU_TRAFFIC_BYTES.get(traffic_key, (None, None))[0],
                         "traffic_source": NCU_TRAFFIC_BYTES.get(traffic_key, (None, None))[1], "peak_source": peak_src, "algorithmic_bytes_per_launch": int(dom_bytes),
                         "ms_per_launch": kernel_ms[dom]},
            "kernels_ms": {k: round(v, 4) for k, v in kernel_ms.items()},
            "composite": comp,
            "final_loss": final_loss,
        }
    else:
        line = None
    if extras:
        # the configuration the other ~8 of 9 training views use (configs/mpv_base.txt:60-68): p=3, pt=3, s=2, alpha=None
        cfg_o = loss_config(margs, ref_view=False)
        ms_o, kms_o, out_o = time_steps(step, lambda: step.step(H, W, ext, intr, res_dev, cfg_o, lr), 3, 1, barrier)
        if rank == 0:
            line["other_view"] = {"loss": "gpnn_lm p=3 pt=3 s=2 alpha=None rou=-2 + rgb/a smooth 0.2 + scale-invariant",
                                  "ms_per_step": ms_o, "steps_per_s": 1000.0 / ms_o, "kernels_ms": kms_o,
                                  "final_loss": float(out_o["loss"])}
    return line


def stage1_step(dev):
    """SURVEY §8(f) N4, second half: one optimisation step of the stage-1 model at configs/mpi_base.txt's shape (D=32,
    36x64 vertices, scale 1.6, 180x320 patch; loop mask, sparsity / smoothness / density terms on) through `MPMesh` +
    `FusedAdam`, next to the oracle port of the same forward + backward on the host (one repetition, 90x160 patch x 4)."""
    from oracle import mpv_oracle as MO
    from videoloop3d_b200 import FusedAdam, MPMesh, default_args_stage1
    H, W = 180, 320
    args = default_args_stage1(d_smooth_loss_weight=0.0, l_smooth_loss_weight=0.05)
    f = 0.8 * W
    intr0 = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float32)
    torch.manual_seed(2)
    m = MPMesh(args, H, W, np.eye(4, dtype=np.float32), intr0, 1.0, 10.0).to(dev).train()
    m.atlas.data[:, 3] = torch.randn(m.atlas.shape[-2:], device=dev) - 1.0
    m.atlas_mask.data.normal_()
    opt = FusedAdam([m.atlas, m.atlas_mask], lr=args.lrate, betas=(0.9, 0.999), eps=1e-8)
    ext, intr = view_for(dict(H=H, W=W))
    gen = torch.Generator(device=dev).manual_seed(3)
    tar = torch.rand((1, 4, H, W), device=dev, generator=gen)
    weights = {k: getattr(args, k + "_loss_weight") for k in ("sparsity", "rgb_smooth", "a_smooth", "density", "l_smooth")}

    def step():
        rgbl, extra = m(H, W, ext, intr)
        loss = ((rgbl - tar) ** 2).mean()
        for k, v in extra.items():
            loss = loss + v.mean() * weights[k]
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / n
    # CPU: the oracle restatement of the same forward + backward on a quarter-size patch (scaled by the pixel ratio)
    hq, wq = H // 2, W // 2
    st, mask = MO.stage1_state(hq, wq, 32, 36, 64, 4, 1.0, 10.0, 1.6, 1.6, seed=2)
    extq, intrq = view_for(dict(H=hq, W=wq))
    torch.set_num_threads(os.cpu_count() or 1)
    a = st.atlas.float().requires_grad_(True)
    am = mask.float().requires_grad_(True)
    t0 = time.perf_counter()
    rgbl_o, extra_o, _ = MO.forward_stage1(st, hq, wq, extq.cpu(), intrq.cpu(), 1.0, 10.0, d_smooth=False, atlas=a, atlas_mask=am,
                                           dtype=torch.float32)
    lo = (rgbl_o ** 2).mean() + sum(extra_o[k] * weights[k] for k in extra_o)
    lo.backward()
    cpu_ms = (time.perf_counter() - t0) * 1e3 * (H * W) / (hq * wq)
    return {"workload": "stage-1 MPMesh step (configs/mpi_base.txt: D=32, 36x64 vertices, scale 1.6, 180x320 patch, loop mask, "
                        "sparsity / rgb_smooth / a_smooth / density / l_smooth) through MPMesh.forward + backward + FusedAdam",
            "ms_per_step": gpu_ms, "steps_per_s": 1000.0 / gpu_ms, "final_loss": float(loss),
            "cpu_port_ms_per_step": cpu_ms, "cpu_sample": f"oracle forward + backward of one {hq}x{wq} patch x {H * W // (hq * wq)}, "
                                                            f"{os.cpu_count() or 1} threads (no Adam)"}


def measure_extras(args, world, rank, dev, group, barrier):
    """Further workloads of the same path (all ranks take part in the sharded ones; the rest is N = 1 only)."""
    from videoloop3d_b200 import FusedLoopStep
    from videoloop3d_b200.train_step import loss_config
    more = {}
    # ---- tile-culled model with static + dynamic tiles: per-thread loads, static-gradient all-reduce
    wl = WORKLOADS["sparse720p"]
    T = wl["T"]
    t0, t1 = (T * rank) // world, (T * (rank + 1)) // world
    model = build_model(wl, dev, t1 - t0, seed=2, first_frame=t0)
    cfg = loss_config(model.args, ref_view=True)
    ext, intr = view_for(wl)
    res = make_target(wl, dev)
    lr = model.args.lrate * 0.01
    step = FusedLoopStep(model, group=group, global_frames=T, timers=True, fused=args.fused, exchange=args.exchange,
                         gather_nn=False)
    H, W = wl["H"], wl["W"]
    ms_s, kms_s, out_s = time_steps(step, lambda: step.step(H, W, ext, intr, res, cfg, lr), 5, 2, barrier)
    pack = model.mesh_pack()
    more["sparse"] = {"workload": "sparse720p: Bernoulli(0.25) tile occupancy, half static / half dynamic, 32x32-texel tiles, "
                                  "D=32, T=48, 720x1280, F=258, ref-view loss", "static_tiles": pack.n_static,
                      "dynamic_tiles": pack.n_dynamic, "atlas_dyn": list(model.atlas_dyn.shape), "atlas": list(model.atlas.shape),
                      "ms_per_step": ms_s, "steps_per_s": 1000.0 / ms_s, "kernels_ms": kms_s, "final_loss": float(out_s["loss"])}
    del step, model, res
    torch.cuda.empty_cache()
    if world > 1:
        return more
    # ---- the reference's own step shape (configs/mpv_base.txt:21-24): host overhead shows here
    wl = WORKLOADS["patch180"]
    model = build_model(wl, dev, wl["T"], seed=2)
    cfg = loss_config(model.args, ref_view=True)
    ext, intr = view_for(wl)
    res = make_target(wl, dev)
    step = FusedLoopStep(model, timers=False, fused=args.fused)
    H, W = wl["H"], wl["W"]
    ms_p, _, out_p = time_steps(step, lambda: step.step(H, W, ext, intr, res, cfg, lr), 30, 5, barrier)
    t_host = time.perf_counter()
    for _ in range(30):
        step.step(H, W, ext, intr, res, cfg, lr)
    host_ms = (time.perf_counter() - t_host) / 30 * 1e3               # launch-side time of a step (no sync inside)
    torch.cuda.synchronize()
    more["patch180"] = {"workload": "dense 180x320 (the reference's step shape), D=32, T=48, F=258, ref-view loss",
                        "ms_per_step": ms_p, "steps_per_s": 1000.0 / ms_p, "host_ms_per_step": host_ms,
                        "final_loss": float(out_p["loss"])}
    del step, model, res
    torch.cuda.empty_cache()
    # ---- BASELINE config 5 sweep and config 0
    more["loss_sweep_ms"] = loss_sweep(dev)
    more["loss_sweep_ms"].update(loss_sweep(dev, extents=((720, 1280),), Ts=(48,), n2s=(256,)))
    torch.cuda.empty_cache()
    more["config0"] = config0_static_render(dev)
    torch.cuda.empty_cache()
    more["stage1"] = stage1_step(dev)
    torch.cuda.empty_cache()
    # ---- the reference's torch operator sequence on this same GPU (the north_star's ">= 10x" denominator)
    if not args.no_gpu_reference:
        torch.cuda.empty_cache()
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        sampler = ClockSampler(0)
        sampler.start()
        try:
            ms_pt, ms_ad, n, (ph, pw) = gpu_torch_reference_step(WORKLOADS["step720p"], dev, 3)
            more["gpu_reference"] = {"ms_per_step": ms_pt + ms_ad, "steps_per_s": 1000.0 / (ms_pt + ms_ad), "repetitions": 3,
                                     "parts_ms": {"patches_fwd_bwd": ms_pt, "adam": ms_ad}, "clocks": sampler.stop(),
                                     "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                                     "what": f"the reference's torch operator sequence (oracle/torch_ref_ops.py: grid_sample / "
                                             f"masked_scatter / cumprod / unfold / bmm / index_add, allow_tf32=False) on this GPU: "
                                             f"{n} patches of {ph}x{pw} covering 720x1280, gradients accumulated, one "
                                             f"torch.optim.Adam step; host rasteriser stand-in not timed"}
        except torch.cuda.OutOfMemoryError as ex:
            sampler.stop()
            more["gpu_reference"] = {"unavailable": f"out of memory: {str(ex)[:100]}"}
    return more


class _OnlyJsonOnStdout:
    """Libraries (NCCL's version banner, torchrun notices) write to file descriptor 1; the contract is ONE JSON line on
    stdout.  Inside this context fd 1 points at stderr; `emit` writes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self._real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._real, 1)
        os.close(self._real)
        return False


_OUT = None


def emit_line(obj):
    text = json.dumps(obj)
    if _OUT is not None:
        _OUT.emit(text)
    else:
        print(text, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="step720p", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: host CPU (the reference arm) or the reference's torch operators on cuda:0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference-operators-on-GPU leg")
    ap.add_argument("--quick", action="store_true", help="headline workload only (no sparse / patch180 / sweep / config0 legs)")
    ap.add_argument("--no-smooth", action="store_true", help="tuning aid: drop the smoothness regularisers")
    ap.add_argument("--exchange", default=None, choices=["p2p", "nccl", "auto"],
                    help="N > 1: exchanges of the band-sharded loss (default: peer-memory stores when available)")
    ap.add_argument("--fused", default=None, choices=["off", "generic", "band", "band-zero", "own", "auto"],
                    help="backward + Adam: separate kernels or one persistent kernel (default: VL3D_FUSED, else auto)")
    args = ap.parse_args()
    global _OUT
    with _OnlyJsonOnStdout() as out:
        _OUT = out
        try:
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
        finally:
            _OUT = None


if __name__ == "__main__":
    main()
