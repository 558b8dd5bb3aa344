#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) as one CSV row per profiled launch:
duration, DRAM bytes and achieved GB/s, DRAM / L2 / tensor-pipe / XU / issue utilisation, registers, grid.
  python tools/ncu_table.py profiles/x.ncu-rep > profiles/x.csv"""
import csv, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_tc_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    out = csv.writer(sys.stdout)
    names = list(WANT.values())
    out.writerow(["id", "kernel"] + names + ["hbm_GBps"])
    for r in rows[2:]:
        rec = {}
        for m, short in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                u = units[col[m]]
                if short == "us":
                    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
                if short.endswith("_MB"):
                    v = v / 1e6 if u == "byte" else v / 1e3 if u == "Kbyte" else v * 1e3 if u == "Gbyte" else v
                rec[short] = v
        name = r[col["Kernel Name"]] if "Kernel Name" in col else "?"
        gb = (rec.get("dram_rd_MB", 0) + rec.get("dram_wr_MB", 0)) / 1e3 / (rec["us"] / 1e6) if rec.get("us") else 0.0
        out.writerow([r[col["ID"]], name[:70]] + [f"{rec[n]:.2f}" if n in rec else "" for n in names] + [f"{gb:.0f}"])


if __name__ == "__main__":
    main(sys.argv[1])
