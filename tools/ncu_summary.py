#!/usr/bin/env python3
"""Text summary of an .ncu-rep for profiles/: key section metrics, pipe utilisation, stall reasons
and the hottest CUDA source lines of the first captured kernel.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r1_xyz.txt
    python tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/r1_launches.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Duration", "Elapsed Cycles", "SM Frequency", "Registers Per Thread", "Block Size", "Grid Size",
        "Shared Memory Configuration Size", "Theoretical Occupancy", "Achieved Occupancy", "Executed Instructions",
        "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Warp Cycles Per Issued Instruction",
        "Avg. Active Threads Per Warp", "Branch Efficiency", "Compute (SM) Throughput", "Memory Throughput",
        "L1/TEX Cache Throughput", "L2 Cache Throughput", "DRAM Throughput", "L1/TEX Hit Rate", "L2 Hit Rate"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
       "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_fma_realtime.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = r[ik].split("(")[0][:70]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(a[1] for a in agg.values())
    print("ncu launch list (gpu__time_duration.sum, --clock-control none): cold-cache, serialised -- compare SHARES")
    print("%-72s %6s %12s %7s" % ("kernel", "n", "total ms", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %12.3f %6.1f%%" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot))
    print("%-72s %6d %12.3f" % ("TOTAL", sum(a[0] for a in agg.values()), tot / 1e6))


def main():
    if sys.argv[1] == "--launches":
        return launches(sys.argv[2])
    rep = sys.argv[1]
    rows = list(csv.reader(run(["-i", rep, "--page", "details", "--csv"]).splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    ids = sorted({r[idx["ID"]] for r in rows[1:]}, key=int)
    for kid in ids:
        name = next(r[idx["Kernel Name"]] for r in rows[1:] if r[idx["ID"]] == kid)
        print("=== launch %s: %s" % (kid, name[:110]))
        for r in rows[1:]:
            if r[idx["ID"]] == kid and r[idx["Metric Name"]] in KEYS:
                print("  %-40s %14s %s" % (r[idx["Metric Name"]], r[idx["Metric Value"]], r[idx["Metric Unit"]]))
    raw = list(csv.reader(run(["-i", rep, "--page", "raw", "--csv"]).splitlines()))
    rh = raw[0]
    for k, row in enumerate(raw[2:]):
        print("--- raw metrics, launch %d" % k)
        for i, h in enumerate(rh):
            if any(h.endswith(x) for x in RAW):
                print("  %-90s %s" % (h, row[i]))
    # source page of the first launch
    out = run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"])
    srows = list(csv.reader(out.splitlines()))
    agg = collections.defaultdict(lambda: [0, 0, 0, ""])
    stalls = collections.Counter()
    cur_file = cur_line = cur_text = None
    shdr = None
    seen = set()
    for r in srows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            if cur_file in seen:
                break
            seen.add(cur_file)
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            shdr = r
            ii, it, isamp = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed"), shdr.index("# Samples")
            continue
        if shdr is None:
            continue
        if r[0] != "":
            cur_line, cur_text = r[0], r[1].strip()
            continue
        try:
            a = agg[(cur_file, int(cur_line))]
            a[0] += int(r[ii]); a[1] += int(r[it]); a[2] += int(r[isamp]); a[3] = cur_text
            for i, h in enumerate(shdr):
                if h.startswith("stall_") and "Not Issued" not in h:
                    stalls[h] += int(r[i])
        except (ValueError, TypeError):
            pass
    ti = sum(a[0] for a in agg.values()) or 1
    tt = sum(a[1] for a in agg.values())
    ts = sum(a[2] for a in agg.values()) or 1
    print("--- source view of launch 0: %d warp instructions, %.2f active threads per instruction" % (ti, tt / ti))
    tot_st = sum(stalls.values()) or 1
    print("stall reasons (all samples): " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100 * v / tot_st)
                                                        for k, v in stalls.most_common(7)))
    print("%-24s %6s %6s %6s  %s" % ("file:line", "inst%", "samp%", "thr", "source"))
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
        print("%-24s %6.2f %6.2f %6.1f  %s" % ("%s:%d" % (f, l), 100 * a[0] / ti, 100 * a[2] / ts, a[1] / max(a[0], 1), a[3][:100]))


if __name__ == "__main__":
    main()
