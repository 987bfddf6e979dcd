#!/bin/bash
# gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_p2p_check.sh r03c 2'   — sharded == single GPU with both exchange modes, then timing
tag=${1:-p2p}; n=${2:-2}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale"; exit 9; }
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    tests/check_sharded.py > $out/${tag}_check_sharded_n$n.log 2>&1
echo "check_sharded rc=$?"
grep -v "^W\|^\*" $out/${tag}_check_sharded_n$n.log | tail -12
for ex in p2p nccl; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 10 --warmup 3 --quick --no-cpu-baseline --no-gpu-reference --exchange $ex > $out/${tag}_bench_n${n}_$ex.json 2> $out/${tag}_bench_n${n}_$ex.err
  echo "bench $ex rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_n${n}_$ex.json").read().strip().splitlines()[-1])
    print("$ex", round(d["value"], 2), "steps/s;", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 2), d["kernels_ms"], d["config"].get("exchange"), "loss", d["final_loss"])
except Exception as e:
    print("$ex: no result", e)
PY
  tail -3 $out/${tag}_bench_n${n}_$ex.err
done
