"""Summarises ncu outputs into small text files for profiles/ (run in the build container, no GPU needed).

  python scripts/ncu_summary.py launches <launches.csv> <out.md>     per-kernel time share of a launch list
  python scripts/ncu_summary.py full <report.ncu-rep> <out.md>       key metrics of a --set full capture
  python scripts/ncu_summary.py facts <streams_per_launch> <out.json> <report.ncu-rep>...
        per bench.py kernel category: dram bytes and FP64-pipe warp instructions per launch (what bench.py's
        roofline.traffic / roofline.fp64 are taken from), averaged over the captured launches of that category
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


def raw_csv(path):
    """--page raw --csv of a report; a path ending in .csv is taken as that export already (made on the GPU box when the
    .ncu-rep itself is too large to bring back)."""
    if path.endswith(".csv"):
        return open(path).read()
    return subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout


def full(path, out):
    txt = raw_csv(path)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none: {path}\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")


# category, kernel-name regex, grid_dim_x of the launch (k_rows6: rows / 8 -> 256 for a 2048-row z level, 384 for the 3072-row x level)
CATS = [("rows_z", r"k_rows6<", 256), ("rows_x", r"k_rows6<", 384), ("pop_z", r"k_pop6", None),
        ("push_z", r"k_push_pairs", None), ("conv_dense3x3", r"k_conv_tc(_2sm)?\(", "ms<1.1"), ("conv_dense5x5", r"k_conv_tc(_2sm)?\(", "ms>=1.1"),
        ("conv_head", r"k_conv_tc_head", None), ("conv_in", r"k_conv_tc_h", None)]


def facts(streams, out, paths):
    import json
    import os
    import re
    acc = {}
    for path in paths:
        only = None
        if ":" in path:                                   # report.ncu-rep:cat1,cat2 -> take only these categories from it
            path, only = path.split(":", 1)
            only = set(only.split(","))
        label = "ncu --set full --clock-control none, " + os.path.basename(path) + " (summaries: profiles/r2_ncu_final*.md)"
        txt = raw_csv(path)
        rows = list(csv.reader(txt.splitlines()))
        hdr = rows[0]

        def col(r, name, default=0.0):
            return float(r[hdr.index(name)].replace(",", "")) if name in hdr and r[hdr.index(name)] else default
        units = rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            gx = int(col(r, "launch__grid_dim_x"))
            ms_ = col(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(
                units[hdr.index("gpu__time_duration.sum")].replace("second", "s").replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns"), 1.0)
            for cat, rx, want_gx in CATS:
                if not re.search(rx, name) or (only is not None and cat not in only):
                    continue
                if isinstance(want_gx, int) and gx != want_gx:
                    continue
                if isinstance(want_gx, str) and not eval(want_gx, {"ms": ms_ * 1024.0 / float(streams)}):   # (the dense 3x3 / 5x5 convs share a kernel)
                    continue
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                rd = col(r, "dram__bytes_read.sum") * scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
                wr = col(r, "dram__bytes_write.sum") * scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
                if "sm__inst_executed_pipe_fp64.sum" in hdr:
                    f64 = col(r, "sm__inst_executed_pipe_fp64.sum")
                    how = "sm__inst_executed_pipe_fp64.sum"
                else:     # 2 FP64 warp instructions per SM and cycle at peak (4 sub-partitions x 16 lanes)
                    f64 = col(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") / 100.0 * 2.0 * col(r, "sm__cycles_active.avg") * 148
                    how = "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active x 2/cycle/SM x sm__cycles_active.avg x 148 SMs"
                a = acc.setdefault(cat, {"n": 0, "dram": 0.0, "f64": 0.0, "ms": 0.0, "inst": 0.0, "src": set(), "how": how})
                a["n"] += 1; a["dram"] += rd + wr; a["f64"] += f64
                a["ms"] += col(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")].replace("second", "s").replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns"), 1.0)
                a["inst"] += col(r, "smsp__inst_executed.sum")
                a["src"].add(label)
                break
    res = {}
    for cat, a in acc.items():
        res[cat] = {"streams_per_launch": int(streams), "dram_bytes_per_launch": a["dram"] / a["n"], "fp64_inst_per_launch": a["f64"] / a["n"],
                    "warp_inst_per_launch": a["inst"] / a["n"], "ncu_ms_per_launch": a["ms"] / a["n"], "launches_captured": a["n"],
                    "fp64_count_from": a["how"], "source": ", ".join(sorted(a["src"]))}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "facts":
        facts(sys.argv[2], sys.argv[3], sys.argv[4:])
        sys.exit(0)
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
