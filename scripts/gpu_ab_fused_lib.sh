#!/bin/bash
# A/B of a library variant on the fused pass: dense step720p (tune_fused) and the tile-culled workload
V=$1
mkdir -p gpurun_out
echo "== dense, default"; timeout 300 python scripts/tune_fused.py --steps 4 --configs "generic:3" 2>&1 | grep bwd_adam_ms
echo "== dense, $V"; VL3D_LIB=$V timeout 300 python scripts/tune_fused.py --steps 4 --configs "generic:3" 2>&1 | grep bwd_adam_ms
echo "== dense, default (again)"; timeout 300 python scripts/tune_fused.py --steps 4 --configs "generic:3" 2>&1 | grep bwd_adam_ms
echo "== dense, $V (again)"; VL3D_LIB=$V timeout 300 python scripts/tune_fused.py --steps 4 --configs "generic:3" 2>&1 | grep bwd_adam_ms
bash scripts/gpu_ab_sparse.sh $V
