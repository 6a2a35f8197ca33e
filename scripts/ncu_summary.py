"""Summarises ncu outputs into small text files for profiles/ (run in the build container, no GPU needed).

  python scripts/ncu_summary.py launches <launches.csv> <out.md>     per-kernel time share of a launch list
  python scripts/ncu_summary.py full <report.ncu-rep> <out.md>       key metrics of a --set full capture
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(path, out):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1e-6) * v
        name = r[ki].split("(")[0].replace("void ", "")
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list: {path}\n\n{sum(cnt.values())} launches, {total:.1f} ms device time "
                "(ncu-serialised, cold-cache: compare SHARES, not absolutes)\n\n| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
        for k in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"| `{k}` | {cnt[k]} | {tot[k]:.2f} | {tot[k]/cnt[k]:.3f} | {100*tot[k]/total:.1f}% |\n")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {path}\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
