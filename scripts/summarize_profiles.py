"""Turn ncu output brought back in gpurun_out/ into small tracked summaries under profiles/.

    python scripts/summarize_profiles.py launches <launches.csv> <out.md>     # per-kernel time shares
    python scripts/summarize_profiles.py full <report.ncu-rep> <out.md>       # key metrics per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ui], 1.0)
        agg.setdefault(r[ki].split("(")[0][:80], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path}): gpu__time_duration.sum per kernel, cold-cache / serialised\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v):.3f} | {100 * sum(v) / tot:.1f}% |\n")
    print(open(out).read())


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n")
        for r in rows[2:]:
            f.write(f"\n## `{r[H.index('Kernel Name')][:90]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in H:
                    f.write(f"| {k} | {r[H.index(k)]} | {units[H.index(k)]} |\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
