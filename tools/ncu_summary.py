#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.txt>     per-kernel device-time shares
  python tools/ncu_summary.py full <file.ncu-rep> <out.txt>         key counters of every captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if r[ui] in ("ns", "nsecond"):
            v /= 1e3
        elif r[ui] in ("ms", "msecond"):
            v *= 1e3
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (%s)\n" % path)
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes.  total %.3f ms over %d launches\n"
                % (tot / 1e3, sum(a[0] for a in agg.values())))
        f.write("%-70s %8s %12s %12s %7s\n" % ("kernel", "launches", "total_us", "avg_us", "share"))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-70s %8d %12.1f %12.2f %6.1f%%\n" % (k[:70], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# key counters from `ncu --set full --clock-control none` capture %s\n" % path)
        for r in rows[2:]:
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write("%-85s %s %s\n" % (k, r[i][:90], units[i]))
            f.write("-" * 60 + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
