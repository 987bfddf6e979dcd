#!/bin/bash
# One gpurun call for the owner mode of the fused backward + Adam kernel: parity tests, step time, DRAM traffic.
#   gpurun --timeout 900 -- 'bash scripts/gpu_own_check.sh r02t'
tag=${1:-own}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale: rebuild before gpurun"; exit 9; }
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q ${PYTEST_K:+-k "$PYTEST_K"} > $out/${tag}_pytest_fused.log 2>&1
echo "pytest fused rc=$?" > $out/${tag}_rc.log
tail -15 $out/${tag}_pytest_fused.log
for mode in ${MODES:-own generic}; do
  timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --quick --fused $mode > $out/${tag}_bench_${mode}.json 2> $out/${tag}_bench_${mode}.err
  echo "bench $mode rc=$?" >> $out/${tag}_rc.log
  tail -3 $out/${tag}_bench_${mode}.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_${mode}.json").read().strip().splitlines()[-1])
    print("$mode", round(d["ms_per_step"], 2), "ms/step; e2e", round(d["e2e"]["ms_per_step"], 2), d["kernels_ms"], "loss", d["final_loss"])
except Exception as e:
    print("$mode: no result", e)
PY
done
for mode in ${NCU_MODES:-own}; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed --clock-control none \
      -k regex:fused_bwd_adam -c 2 --csv --log-file $out/${tag}_ncu_fused_${mode}.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference --quick --fused $mode > $out/${tag}_ncu_fused_${mode}.log 2>&1
  echo "ncu $mode rc=$?" >> $out/${tag}_rc.log
  grep -E "fused_bwd_adam" $out/${tag}_ncu_fused_${mode}.csv | cut -d, -f5,12- | head -14
done
cat $out/${tag}_rc.log
