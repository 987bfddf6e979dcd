#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_profile_r02.sh r03p'
# Evidence kept under profiles/: launch list of the default bench command, full-set captures of the hot kernels.
tag=${1:-r03p}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale"; exit 9; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
# every launch of the step (cold-cache, serialised: shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_step720p.csv \
    python bench.py --steps 2 --warmup 1 --quick --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
# DRAM traffic of the fused kernel at 720p (single-pass metrics: no replay of the 90 GB state)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum \
    --clock-control none -k regex:fused_bwd_adam -s 1 -c 1 --csv --log-file $out/${tag}_fused_720p.csv \
    python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline > $out/${tag}_fused_720p.log 2>&1
echo "ncu fused 720p rc=$?"
# full sets (with source) at 360p, where the replay state fits: fused backward + Adam, search, vote, render
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fused_bwd_adam|patchnn_strip8|vote_loss|composite_render_tma' \
    -s 4 -c 4 -o $out/${tag}_full_step360p -f python bench.py --workload step360p --steps 1 --warmup 1 --quick --no-cpu-baseline > $out/${tag}_full_step360p.log 2>&1
echo "ncu full 360p rc=$?"
ls -la $out | grep ${tag}
