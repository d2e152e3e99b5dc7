"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<round>_launches.md
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/<round>_cover_eval_full.md
"""
import collections
import csv
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg.per_second",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def launches(path):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1.0)
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:110]}` | {n} | {t:.3f} | {100 * t / total:.2f} % |")
    print(f"\ntotal {total:.3f} ms over {sum(a[0] for a in agg.values())} launches "
          "(ncu-serialised, cold-cache: compare shares, not absolutes)")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"### {name[:120]}\n\n| metric | unit | value |\n|---|---|---:|")
        for m in KEY_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {units[i]} | {vals[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
