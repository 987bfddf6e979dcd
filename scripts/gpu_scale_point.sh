#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_scale_point.sh r03e N'   — the default bench line at N GPUs (+ NCCL-exchange A/B, quick)
tag=${1:-scale}; n=${2:-8}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale"; exit 9; }
run() {  # name, extra args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 10 --warmup 3 --no-gpu-reference --no-cpu-baseline $2 > $out/${tag}_bench_n${n}_$1.json 2> $out/${tag}_bench_n${n}_$1.err
  echo "bench $1 rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_n${n}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), d["kernels_ms"], d["config"].get("exchange"), "loss", d["final_loss"])
    for k in ("sharded_check", "sparse"):
        if k in d: print(k, json.dumps(d[k])[:400])
except Exception as e:
    print("$1: no result", e)
PY
}
run p2p ""
run nccl "--quick --exchange nccl"
