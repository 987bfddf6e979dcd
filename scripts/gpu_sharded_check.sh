#!/bin/bash
# gpurun --gpus N --timeout 420 -- 'bash scripts/gpu_sharded_check.sh r01d N'
tag=${1:-check}; n=${2:-2}
out=gpurun_out
mkdir -p $out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    tests/check_sharded.py > $out/${tag}_check_sharded_n$n.log 2>&1
echo "check_sharded rc=$?" > $out/${tag}_rc_n$n.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
echo "bench rc=$?" >> $out/${tag}_rc_n$n.log
cat $out/${tag}_rc_n$n.log
grep -v "^W\|^\*" $out/${tag}_check_sharded_n$n.log | tail -8
tail -1 $out/${tag}_bench_n$n.json
