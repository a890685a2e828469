#!/usr/bin/env python
"""Top source lines of the first launch in an .ncu-rep by executed warp instructions.
Usage: python tools/ncu_lines.py prof.ncu-rep [warps] [top]"""
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


def main(rep, warps=1, top=45):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    cur, H, out = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            H = r
            continue
        if H and cur:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            d = dict(zip(H[4:], r[4:]))
            out.append((cur, ln, num(d.get("Instructions Executed")), num(d.get("# Samples")), r[1][:100]))
    tot = sum(o[2] for o in out) or 1
    ts = sum(o[3] for o in out) or 1
    byfile = {}
    for o in out:
        a = byfile.setdefault(o[0], [0, 0])
        a[0] += o[2]
        a[1] += o[3]
    print(f"warp instructions {tot} ({tot / warps:.0f} per warp), stall samples {ts}")
    for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:28s} inst {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / ts:5.1f}%")
    for o in sorted(out, key=lambda o: -o[2])[:top]:
        print(f"{o[0][:16]:16s}:{o[1]:4d} {o[2] / warps:7.1f}/warp {100 * o[2] / tot:5.1f}% samp {100 * o[3] / ts:4.1f}%  {o[4]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1, int(sys.argv[3]) if len(sys.argv) > 3 else 45)
