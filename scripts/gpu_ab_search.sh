#!/bin/bash
# A/B of a search-kernel variant build against the regular library (one gpurun call): scripts/gpu_ab_search.sh <variant .so>
V=${1:-videoloop3d_b200/lib/libvl3d_rowfree.so}
mkdir -p gpurun_out
for cfg in "--p 11 --s 4" "--p 7 --s 4 --alpha 10000" "--p 15 --s 4" "--p 11 --s 4 --H 180 --W 320 --T 96"; do
  echo "== $cfg"
  timeout 120 python scripts/tune_search.py $cfg --reps 3 --dump /tmp/nn_ref.pt 2>&1 | grep -v Warn | grep "search\|ho="
  VL3D_LIB=$V timeout 120 python scripts/tune_search.py $cfg --reps 3 --check /tmp/nn_ref.pt 2>&1 | grep -v Warn | grep "search\|indices differing from /"
done
