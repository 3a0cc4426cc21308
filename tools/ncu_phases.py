#!/usr/bin/env python
"""Attribute the executed warp instructions and stall samples of one k_tile_pass launch to the
kernel's phases, from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`.
Every SASS address is counted once (inlining lists it under several source lines); the phase is
decided by the source-line RANGES given in a small table at the top (edit when tile.cu moves).
Usage: ncu_phases.py dump.csv [kernel_index=1] [ranges.json]"""
import csv
import sys
import json
import collections


def load(path, want):
    kern = 0
    hdr = None
    cur = None
    fpath = ""
    addr_lines = collections.defaultdict(set)
    info = {}
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
            if fpath.endswith(".cu") or fpath.endswith(".cuh"):
                kern += 1 if fpath.endswith("tile.cu") else 0
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or kern != want:
            continue
        if r[0] != "":
            cur = (fpath.split("/")[-1], int(r[0])) if r[0].isdigit() else None
            continue
        if cur is None:
            continue
        a = r[2]
        addr_lines[a].add(cur)
        if a not in info:
            d = dict(zip(hdr[:2] + ["Address", "Sass"] + hdr[4:], r))
            try:
                info[a] = (d["Sass"], int(d["Instructions Executed"]), int(d["# Samples"]))
            except ValueError:
                continue
    return addr_lines, info


def main():
    path = sys.argv[1]
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ranges = json.load(open(sys.argv[3])) if len(sys.argv) > 3 else None
    addr_lines, info = load(path, want)
    ti = sum(v[1] for v in info.values()) or 1
    ts = sum(v[2] for v in info.values()) or 1
    print(f"{len(info)} SASS instructions, {ti} warp instr executed, {ts} samples")
    agg = collections.defaultdict(lambda: [0, 0])
    for a, (s, i, sm) in info.items():
        lines = sorted(l for f, l in addr_lines[a] if f == "tile.cu")
        b = "other"
        if ranges:
            for name, lo, hi in ranges:          # first match wins
                if any(lo <= l <= hi for l in lines):
                    b = name
                    break
        agg[b][0] += i
        agg[b][1] += sm
    for b, (i, sm) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{b:44s} instr {i / ti * 100:5.1f}%  samples {sm / ts * 100:5.1f}%")


if __name__ == "__main__":
    main()
