#!/bin/bash
# ncu --set full (with source) of the fused backward + Adam kernel on the tile-culled 720p workload
tag=${1:-r04}
out=gpurun_out
mkdir -p $out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fused_bwd_adam -s 1 -c 1 -o $out/${tag}_full_sparse720p_fused -f \
    python bench.py --workload sparse720p --steps 1 --warmup 1 --quick --no-cpu-baseline > $out/${tag}_full_sparse720p.log 2>&1
echo "ncu rc=$?"
tail -3 $out/${tag}_full_sparse720p.log
ls -la $out | grep ${tag}
