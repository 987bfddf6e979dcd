#!/bin/bash
# A/B of a library variant on the tile-culled 720p workload: scripts/gpu_ab_sparse.sh <variant .so> [more variants]
mkdir -p gpurun_out
run() { timeout 300 python bench.py --workload sparse720p --quick --no-cpu-baseline --steps 6 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms_per_step', round(d['ms_per_step'], 3), 'kernels', {k: round(v, 3) for k, v in d['kernels_ms'].items()}, 'loss', d['final_loss'])"; }
run default
for v in "$@"; do VL3D_LIB=$v run $v; done
