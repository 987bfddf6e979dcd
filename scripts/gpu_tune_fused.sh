#!/bin/bash
# gpurun --timeout 1200 -- 'bash scripts/gpu_tune_fused.sh r02b "<configs>"'
tag=${1:-tune}
cfgs=${2:-"off,generic:3,generic:2,band:3,band-zero:3"}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale: rebuild before gpurun"; exit 9; }
timeout 200 python -m pytest tests/test_gpu_fused.py -x -q > $out/${tag}_pytest_fused.log 2>&1
echo "pytest fused rc=$?"; tail -3 $out/${tag}_pytest_fused.log
timeout 500 python scripts/tune_fused.py --steps 3 --configs "$cfgs" > $out/${tag}_tune.jsonl 2> $out/${tag}_tune.err
echo "tune rc=$?"; cat $out/${tag}_tune.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:fused_bwd_adam --csv --log-file $out/${tag}_tune_ncu.csv \
    python scripts/tune_fused.py --steps 1 --configs "$cfgs" > $out/${tag}_tune_ncu.log 2>&1
echo "ncu rc=$?"
python - <<PY
import csv
rows = [r for r in csv.reader(open("$out/${tag}_tune_ncu.csv")) if len(r) > 10 and "fused_bwd_adam" in r[4]]
by = {}
for r in rows:
    by.setdefault(r[0], {})[r[-3]] = float(r[-1].replace(",", ""))
for k, m in by.items():
    print(k, "ms", round(m.get("gpu__time_duration.sum", 0) / 1e6, 2), "R GB", round(m.get("dram__bytes_read.sum", 0) / 1e9, 1), "W GB", round(m.get("dram__bytes_write.sum", 0) / 1e9, 1))
PY
