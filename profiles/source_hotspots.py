#!/usr/bin/env python
"""Top stall-sample SASS instructions of one kernel from `ncu --page source --csv` output.
usage: source_hotspots.py <rep.ncu-rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys


def main(rep, pattern, top_n=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the export holds one block per kernel instance: "Kernel Name" row, header row, data rows
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    b = blocks[0]
    hdr, data = b["hdr"], b["data"]
    i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[i_s]) for r in data)
    print(b["name"][:60], "| samples", tot, "| warp instructions", sum(int(r[i_ex]) for r in data), "| SASS lines", len(data))
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
    print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    top = sorted(enumerate(data), key=lambda t: -int(t[1][i_s]))[:int(top_n)]
    for i, r in sorted(top):
        why = max(stall_cols, key=lambda c: int(r[c] or 0))
        print("%5d %-64s %7s %5.1f%% %-14s exec=%s" % (i, r[i_src].strip()[:64], r[i_s], 100 * int(r[i_s]) / tot, hdr[why], r[i_ex]))


if __name__ == "__main__":
    main(*sys.argv[1:])
