#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` dump per CUDA source
line: share of executed warp instructions and of stall samples (first kernel of the report by
default).  Usage: ncu_lines.py dump.csv [kernel_index] [top_n]"""
import csv
import sys
import collections


def main():
    path = sys.argv[1]
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    kern = 0
    hdr = None
    lines = collections.OrderedDict()
    cur = None
    fpath = ""
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            if fpath.endswith("tile.cu") or True:
                pass
            continue
        if r[0] == "Kernel Name":
            kern += 1
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr[:2] + ["Address", "Sass"] + hdr[4:], r))
        if r[0] != "":
            cur = (fpath.split("/")[-1], r[0], r[1].strip()[:90])
            lines.setdefault(cur, [0, 0, collections.Counter()])
        elif cur is not None:
            try:
                ie = int(d["Instructions Executed"])
                sm = int(d["# Samples"])
            except (KeyError, ValueError):
                continue
            lines[cur][0] += ie
            lines[cur][1] += sm
            lines[cur][2][d["Sass"].split()[0 if not d["Sass"].strip().startswith("@") else 1].split(".")[0]] += ie
    ti = sum(v[0] for v in lines.values()) or 1
    ts = sum(v[1] for v in lines.values()) or 1
    print(f"total warp instr {ti}, samples {ts}")
    for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        ops = ",".join(f"{o}:{c * 100 // max(v[0], 1)}" for o, c in v[2].most_common(4))
        print(f"{k[0]}:{k[1]:>4s} instr {v[0] / ti * 100:5.1f}% samp {v[1] / ts * 100:5.1f}%  {ops:40s} | {k[2]}")


if __name__ == "__main__":
    main()
