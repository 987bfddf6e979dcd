#!/usr/bin/env python
"""Sweep of the fused backward + Adam kernel's schedule knobs at the bench workload (one process, one model).

    python scripts/tune_fused.py [--workload step720p] [--steps 3] [--configs "mode:ctas:row_block:zero_ahead:adam_lag:seg_texels,..."]

Prints one line per configuration: CUDA-event ms of the fused kernel.  Run it under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:fused_bwd_adam
to get the DRAM traffic of every launch (launch order = configuration order x (1 warm-up + steps))."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402

DEFAULT = "off,generic:3,generic:2,band:3,band-zero:3"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="step720p")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--configs", default=DEFAULT)
    args = ap.parse_args()
    from videoloop3d_b200 import FusedLoopStep
    from videoloop3d_b200.train_step import loss_config
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = bench.WORKLOADS[args.workload]
    model = bench.build_model(wl, dev, wl["T"])
    cfg = loss_config(model.args, ref_view=True)
    ext, intr = bench.view_for(wl)
    res = bench.make_target(wl, dev)
    lr = model.args.lrate * 0.01
    step = FusedLoopStep(model, timers=True, fused="off")
    H, W = wl["H"], wl["W"]
    pr = torch.cuda.get_device_properties(0)
    print(json.dumps({"L2": pr.L2_cache_size, "persisting_max": getattr(pr, "persisting_l2_cache_max_size", None),
                      "access_policy_max_window": getattr(pr, "access_policy_max_window_size", None)}), flush=True)
    for spec in args.configs.split(","):
        f = spec.split(":")
        mode = f[0]
        ctas = int(f[1]) if len(f) > 1 else 3
        opts = dict(ctas_per_sm=ctas)
        for name, i in (("row_block", 2), ("zero_ahead", 3), ("adam_lag", 4), ("seg_texels", 5)):   # e.g. generic:3::::65536
            if len(f) > i and f[i] != "":
                opts[name] = int(f[i])
        step.fused, step.fused_opts = mode, opts
        step._sched_cache.clear()
        step.timers.clear()
        for _ in range(1 + args.steps):
            out = step.step(H, W, ext, intr, res, cfg, lr)
        torch.cuda.synchronize()
        ms = step.timer_ms(skip=1)
        key = "fused_bwd_adam" if mode != "off" else None
        t = ms.get("fused_bwd_adam") if key else ms.get("grad_zero", 0) + ms.get("composite_bwd", 0) + ms.get("adam", 0)
        st = step.last_schedule.stats if (step.last_schedule is not None and mode != "off") else {}
        print(json.dumps({"config": spec, "bwd_adam_ms": round(t, 3), "loss": float(out["loss"]), "stats": st}), flush=True)


if __name__ == "__main__":
    main()
