#!/bin/bash
# One gpurun call that re-validates HEAD on a B200 and brings back the evidence kept under profiles/:
#   gpurun --timeout 900 -- 'bash scripts/gpu_round_check.sh r04'
# Every leg has its own timeout; results land in gpurun_out/<tag>_*.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?" > $out/${tag}_rc.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?" >> $out/${tag}_rc.log
timeout 90 python __graft_entry__.py --smoke > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> $out/${tag}_rc.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_step720p.csv \
    python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?" >> $out/${tag}_rc.log
cat $out/${tag}_rc.log
tail -3 $out/${tag}_pytest_gpu.log
tail -2 $out/${tag}_smoke.log
cat $out/${tag}_bench_n1.json
