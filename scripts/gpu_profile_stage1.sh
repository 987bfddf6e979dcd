#!/bin/bash
# ncu captures of the stage-1 step (MPMesh at configs/mpi_base.txt's shape): launch list + full set of the terms kernels
tag=${1:-r04}
out=gpurun_out
mkdir -p $out
cat > /tmp/stage1_step.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import json, torch, bench
print(json.dumps(bench.stage1_step(torch.device("cuda:0"))))
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $out/${tag}_launches_stage1.csv \
    python /tmp/stage1_step.py > $out/${tag}_stage1_under_ncu.log 2>&1
echo "ncu launches rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:composite_terms|composite_fwd_kernel|composite_bwd_kernel' -s 12 -c 5 \
    -o $out/${tag}_full_stage1 -f python /tmp/stage1_step.py > $out/${tag}_full_stage1.log 2>&1
echo "ncu full rc=$?"
ls -la $out | grep ${tag}_.*stage1
