#!/bin/bash
# gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi_check.sh r02j N [pytest]'
tag=${1:-multi}; n=${2:-2}; dotests=${3:-}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale"; exit 9; }
if [ -n "$dotests" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x --tb=short > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest_gpu.log
  grep -E "^E  " $out/${tag}_pytest_gpu.log | cut -c1-250 | head -12
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tests/check_sharded.py > $out/${tag}_check_sharded_n$n.log 2>&1
echo "check_sharded rc=$?"; grep -E "OK|MISMATCH|Error" $out/${tag}_check_sharded_n$n.log | tail -8
for shard in rows frames; do
  VL3D_LOSS_SHARD=$shard timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $n --steps 10 --warmup 3 --no-gpu-reference $( [ $shard = frames ] && echo --quick ) > $out/${tag}_bench_n${n}_$shard.json 2> $out/${tag}_bench_n${n}_$shard.err
  echo "bench $shard rc=$?"; tail -2 $out/${tag}_bench_n${n}_$shard.err | cut -c1-300
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_n${n}_$shard.json").read().strip().splitlines()[-1])
    print("$shard", "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), d["kernels_ms"], "loss", d["final_loss"])
    for k in ("sharded_check", "other_view", "sparse"):
        if k in d: print(k, json.dumps(d[k])[:500])
except Exception as e:
    print("$shard: no result", e)
PY
done
