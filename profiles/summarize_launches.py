#!/usr/bin/env python
"""Per-kernel time / DRAM traffic table of ONE train step from an ncu CSV
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ... python bench.py --steps 1 --warmup 3).
    python profiles/summarize_launches.py gpurun_out/x.csv [step_index] > profiles/x.txt"""
import collections
import csv
import re
import sys


def main(path, step_idx=3):
    lines = [l for l in open(path) if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        i, m = int(row["ID"]), row["Metric Name"]
        v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
        r = recs.setdefault(i, {"name": row["Kernel Name"]})
        if m == "gpu__time_duration.sum":
            r["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            r["rd" if "read" in m else "wr"] = v * mult
    rows = list(recs.values())
    starts = [i for i, r in enumerate(rows) if "cls_prepend_fwd" in r["name"]]
    s, e = starts[step_idx], starts[step_idx + 1]
    step = rows[s:e]
    agg = collections.OrderedDict()
    for r in step:
        n = re.sub(r"\(.*", "", r["name"])
        n = re.sub(r"void |lstc::|<unnamed>::", "", n)
        a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += r["us"]; a[2] += r.get("rd", 0); a[3] += r.get("wr", 0)
    tot = sum(r["us"] for r in step)
    print(f"# source: {path}; one LTN-SHT train step (1280 windows, fwd+bwd, train-mode dropout), device-resident inputs")
    print("# per-launch values summed per kernel; times are serialised / cold-cache: compare SHARES. GB/s = DRAM (rd+wr) / time")
    print(f"# launches in step: {len(step)}   sum of kernel durations: {tot / 1e3:.2f} ms")
    print(f"{'ms':>8} {'share':>6} {'n':>3} {'dram_rd_GB':>10} {'dram_wr_GB':>10} {'GB/s':>7}  kernel")
    for n, (c, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / 1e3:8.3f} {100 * us / tot:5.1f}% {c:3d} {rd / 1e9:10.3f} {wr / 1e9:10.3f} {(rd + wr) / us / 1e3:7.0f}  {n[:100]}")
    g = [r for r in step if "gemm_bf16" in r["name"]]
    print(f"# GEMM launches: {len(g)}, mean DRAM traffic per launch {sum(r['rd'] + r['wr'] for r in g) / len(g) / 1e9:.3f} GB")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
