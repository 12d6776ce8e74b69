#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one line per captured launch with the metrics the roofline
discussion needs (duration, DRAM bytes and throughput, L2 hit rate, occupancy, registers, shared memory, FP64 pipe).
usage: ncu_summary.py raw.csv [> summary.csv]"""
import csv
import sys

WANT = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "duration_us"), ("dram__bytes_read.sum", "dram_read_MB"),
        ("dram__bytes_write.sum", "dram_write_MB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn_B"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cycles_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_scoreboard_pct"),
        ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall_barrier_pct")]


def num(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.lower()
    if u in ("ns", "nsecond"):
        x /= 1e3
    elif u in ("ms", "msecond"):
        x *= 1e3
    elif u in ("s", "second"):
        x *= 1e6
    elif u == "byte" or u == "bytes":
        pass
    elif u == "kbyte":
        x *= 1e3
    elif u == "mbyte":
        x *= 1e6
    elif u == "gbyte":
        x *= 1e9
    return x


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    head, units = rows[hi], rows[hi + 1]
    col = {h: i for i, h in enumerate(head)}
    out = csv.writer(sys.stdout)
    names = [n for k, n in WANT if k in col]
    out.writerow(names + ["dram_GBs"])
    for r in rows[hi + 2:]:
        if len(r) < len(head):
            continue
        vals = {}
        for k, n in WANT:
            if k not in col:
                continue
            v = num(r[col[k]], units[col[k]])
            if n in ("dram_read_MB", "dram_write_MB") and isinstance(v, float):
                v = v / 1e6
            if n == "kernel":
                v = v.split("(")[0][:60]
            vals[n] = v
        gbs = ""
        try:
            gbs = "%.0f" % ((vals["dram_read_MB"] + vals["dram_write_MB"]) * 1e6 / (vals["duration_us"] * 1e-6) / 1e9)
        except Exception:
            pass
        out.writerow([("%.3f" % vals[n] if isinstance(vals[n], float) else vals[n]) for n in names] + [gbs])


if __name__ == "__main__":
    main(sys.argv[1])
