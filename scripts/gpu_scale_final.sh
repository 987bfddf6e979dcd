#!/bin/bash
# gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_scale_final.sh r03u N'   — the default bench command at N GPUs
tag=${1:-final}; n=${2:-8}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale"; exit 9; }
if [ "$n" = "1" ]; then
  timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
else
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
fi
echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2), d["kernels_ms"], d["config"].get("exchange"), "loss", d["final_loss"])
    for k in ("sharded_check",):
        if k in d: print(k, json.dumps(d[k])[:300])
    for k in ("other_view", "sparse", "patch180"):
        if k in d: print(k, d[k].get("ms_per_step"))
except Exception as e:
    print("no result", e)
PY
