#!/bin/bash
# gpurun --timeout 900 -- 'bash scripts/gpu_own_sweep.sh r02x'   — owner mode at 3 / 2 CTAs per SM, parity first
tag=${1:-ownsweep}
out=gpurun_out
mkdir -p $out
python -c "from videoloop3d_b200 import build; import sys; sys.exit(1 if build.needs_build() else 0)" || { echo "libvl3d.so is stale"; exit 9; }
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q > $out/${tag}_pytest_fused.log 2>&1
echo "pytest fused rc=$?"; tail -3 $out/${tag}_pytest_fused.log
for cfg in ${CFGS:-own:3 own:2 generic:3}; do
  mode=${cfg%%:*}; n=${cfg##*:}
  VL3D_FUSED_CTAS_PER_SM=$n timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference --quick --fused $mode > $out/${tag}_bench_${mode}_$n.json 2> $out/${tag}_bench_${mode}_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_bench_${mode}_$n.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["ms_per_step"], 2), "ms/step", d["kernels_ms"], "loss", d["final_loss"])
except Exception as e:
    print("$cfg: no result", e)
PY
done
