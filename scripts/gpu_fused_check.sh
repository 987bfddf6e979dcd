#!/bin/bash
# One gpurun call for the fused backward + Adam kernel: parity tests, the step in every schedule, DRAM traffic (ncu).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_fused_check.sh r02a'
tag=${1:-fused}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale: rebuild before gpurun"; exit 9; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_tma.py -x -q > $out/${tag}_pytest_fused.log 2>&1
echo "pytest fused rc=$?" > $out/${tag}_rc.log
tail -5 $out/${tag}_pytest_fused.log
for mode in ${MODES:-off generic band band-zero}; do
  timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fused $mode > $out/${tag}_bench_${mode}.json 2> $out/${tag}_bench_${mode}.err
  echo "bench $mode rc=$?" >> $out/${tag}_rc.log
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_${mode}.json").read().strip().splitlines()[-1])
    print("$mode", round(d["ms_per_step"], 2), "ms/step; e2e", round(d["e2e"]["ms_per_step"], 2), d["kernels_ms"], "loss", d["final_loss"])
except Exception as e:
    print("$mode: no result", e)
PY
done
for mode in ${NCU_MODES:-band-zero band generic}; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:fused_bwd_adam -c 2 --csv --log-file $out/${tag}_ncu_fused_${mode}.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --fused $mode > $out/${tag}_ncu_fused_${mode}.log 2>&1
  echo "ncu $mode rc=$?" >> $out/${tag}_rc.log
  grep -E "fused_bwd_adam" $out/${tag}_ncu_fused_${mode}.csv | cut -d, -f5,12- | head -8
done
timeout 420 python -m pytest tests -m gpu -x -q --durations=8 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest all rc=$?" >> $out/${tag}_rc.log
tail -4 $out/${tag}_pytest_gpu.log
cat $out/${tag}_rc.log
