#!/usr/bin/env python
"""Top stall sites of one launch from an .ncu-rep captured with --import-source on.
    python profiles/sass_hot.py gpurun_out/x.ncu-rep [launch index] [top N]
Prints the SASS instructions with the most warp-stall samples, the dominant stall reason of each, and the
executed-instruction count, so that a latency-bound kernel's waiting points can be read off without a GUI."""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(float(x.replace(",", ""))) if x else 0
    except ValueError:
        return 0


def main(path, launch=0, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip",
                          str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1][:150])
    hdr = rows[1]
    si = hdr.index("# Samples")
    ii = hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break  # next launch
        if len(r) == len(hdr) and r != hdr:
            body.append(r)
    tot = sum(num(r[si]) for r in body)
    tot_inst = sum(num(r[ii]) for r in body)
    print(f"total samples {tot}, warp instructions executed {tot_inst}")
    agg = {}
    for i, h in stall_cols:
        agg[h] = sum(num(r[i]) for r in body)
    print("stall totals:", ", ".join(f"{h[6:]}={v}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    ranked = sorted(range(len(body)), key=lambda k: -num(body[k][si]))[:top]
    for k in sorted(ranked):
        r = body[k]
        reasons = sorted(((num(r[i]), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        print(f"{k:5d} {num(r[si]):7d} ({100.0 * num(r[si]) / max(tot, 1):4.1f}%) inst={r[ii]:>8s} "
              f"{reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}  {r[1][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
