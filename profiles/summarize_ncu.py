#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into a compact per-launch table of the metrics DESIGN.md / bench.py quote.
    python profiles/summarize_ncu.py gpurun_out/x.ncu-rep [more ...] > profiles/x_summary.txt
A `.csv` argument is taken as the output of `ncu -i x.ncu-rep --page raw --csv` (made on the GPU box, where the report
itself is too large to bring back).
"""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("launch__occupancy_limit_registers", "occ_regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("smsp__cycles_active.avg", "smsp_active_cyc"),
        ("sm__cycles_elapsed.max", "elapsed_cyc"), ("lts__t_bytes.sum", "l2_bytes")]


def main(path):
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    if len(rows) < 3:
        print("# " + path + ": no kernel rows")
        return
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), n, units[hdr.index(m)]) for m, n in WANT if m in hdr]
    ki = hdr.index("Kernel Name")
    print("# " + path)
    print("kernel | " + " | ".join(f"{n} [{u}]" for _, n, u in cols))
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("lstc::", "")
        vals = []
        for i, _, _ in cols:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(name[:60] + " | " + " | ".join(vals))


if __name__ == "__main__":
    for a in sys.argv[1:]:
        main(a)
