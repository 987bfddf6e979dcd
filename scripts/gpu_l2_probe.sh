#!/bin/bash
# gpurun --timeout 600 -- 'bash scripts/gpu_l2_probe.sh r02d'
tag=${1:-l2}
out=gpurun_out
mkdir -p $out
for cfg in "32 512 0" "32 128 0" "32 512 64" "64 1024 96"; do
  set -- $cfg
  name=${tag}_l2probe_g$1_s$2_p$3
  timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none --csv \
      --log-file $out/$name.csv ./build_probe/l2_probe $1 $2 $3 > $out/$name.log 2>&1
  echo "== G=$1 MB, stream=$2 MB, persist=$3 MB"; head -3 $out/$name.log
  python - <<PY
import csv
rows = [r for r in csv.reader(open("$out/$name.csv")) if len(r) > 10 and r[0].isdigit()]
by = {}
for r in rows:
    by.setdefault(int(r[0]), {"k": r[4]})[r[-3]] = float(r[-1].replace(",", ""))
ids = sorted(by)
names = ["normal", "stream .cs", "stream evict_first", "stream hint evict_first", "zero evict_last + stream evict_first",
         "zero/red hint evict_last + stream evict_first", "zero/red evict_last, stream normal", "zero evict_last, stream ef, scalar atomics, read ef"]
for e in range(len(ids) // 6):
    ph = [by[ids[e * 6 + i]] for i in range(6)]
    f = lambda m: "%6.1f R %6.1f W" % (m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6)
    print("%-55s zero[%s] stream[%s] red[%s] read[%s]" % (names[e], f(ph[1]), f(ph[2]), f(ph[3]), f(ph[5])))
PY
done
