#!/usr/bin/env python3
"""Summarise an .ncu-rep per CUDA source line: instructions executed, average active threads,
stall samples.  Usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [--kernel NAME_SUBSTRING]
(the first captured launch whose kernel name contains the substring; default: the first launch)"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 40
    extra = []
    if "--kernel" in sys.argv:
        extra = ["--kernel-name", "regex:" + sys.argv[sys.argv.index("--kernel") + 1], "--launch-count", "1"]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"] + extra,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = defaultdict(lambda: [0, 0, 0, ""])  # inst, thread inst, samples, text
    cur_file, cur_line, cur_text, hdr = "", None, "", None
    seen_files = set()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            if cur_file in seen_files:  # second captured launch starts: keep the first only
                break
            seen_files.add(cur_file)
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst = hdr.index("Instructions Executed")
            i_tinst = hdr.index("Thread Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            cur_line, cur_text = r[0], r[1].strip()
            continue
        try:
            key = (cur_file, int(cur_line))
            a = agg[key]
            a[0] += int(r[i_inst]); a[1] += int(r[i_tinst]); a[2] += int(r[i_samp]); a[3] = cur_text
        except (ValueError, TypeError):
            pass
    tot_inst = sum(a[0] for a in agg.values()) or 1
    tot_t = sum(a[1] for a in agg.values())
    tot_s = sum(a[2] for a in agg.values()) or 1
    print("total warp-inst %d, avg active threads %.2f, samples %d" % (tot_inst, tot_t / tot_inst, tot_s))
    print("%-22s %6s %7s %7s %6s  %s" % ("file:line", "inst%", "cum%", "samp%", "thr", "source"))
    cum = 0.0
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        cum += 100.0 * a[0] / tot_inst
        print("%-22s %6.2f %7.2f %7.2f %6.1f  %s" % ("%s:%d" % (f, l), 100.0 * a[0] / tot_inst, cum, 100.0 * a[2] / tot_s,
                                                     a[1] / max(a[0], 1), a[3][:90]))
    # per-file-region summary
    reg = defaultdict(lambda: [0, 0, 0])
    for (f, l), a in agg.items():
        reg[f][0] += a[0]; reg[f][1] += a[1]; reg[f][2] += a[2]
    for f, a in sorted(reg.items(), key=lambda kv: -kv[1][0]):
        print("FILE %-22s inst %5.1f%% thr %.1f samp %5.1f%%" % (f, 100.0 * a[0] / tot_inst, a[1] / max(a[0], 1), 100.0 * a[2] / tot_s))


if __name__ == "__main__":
    main()
