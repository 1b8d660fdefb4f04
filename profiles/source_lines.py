#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of one kernel (`ncu --import-source on` capture).
usage: source_lines.py <rep.ncu-rep> <kernel regex> [top N]"""
import collections
import csv
import subprocess
import sys


def main(rep, pattern, top_n=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern,
                          "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    hdr, data = None, []
    for r in csv.reader(out.splitlines()):
        if hdr is None and "Source" in r and "Instructions Executed" in r:
            hdr = r
        elif hdr is not None and len(r) == len(hdr):
            if r[0] == "Line No":  # second kernel instance
                break
            data.append(r)
    i_e, i_n = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = collections.OrderedDict()
    cur = None
    for r in data:
        if r[0].strip():  # a CUDA source line; its SASS rows follow
            cur = r[0].strip()
            agg.setdefault(cur, [0, 0, r[1].strip()[:100]])
        elif cur:
            try:
                agg[cur][0] += int(r[i_e])
                agg[cur][1] += int(r[i_n])
            except ValueError:
                pass
    tot_e = sum(v[0] for v in agg.values()) or 1
    tot_n = sum(v[1] for v in agg.values()) or 1
    print("warp instructions %d, samples %d" % (tot_e, tot_n))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(top_n)]:
        print("%6s  instr %5.1f%%  samples %5.1f%%  %s" % (k, 100.0 * v[0] / tot_e, 100.0 * v[1] / tot_n, v[2]))


if __name__ == "__main__":
    main(*sys.argv[1:])
