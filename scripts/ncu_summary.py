#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few numbers the design is judged on.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")], " grid", r[hdr.index("Grid Size")], " block", r[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            print(f"  {'traffic = dram read + write (units as above)':75s} {rd + wr:18.3f}")
        except Exception:
            pass
        print("  stall reasons (warps per issue-active cycle; > 0.15 shown):")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if v > 0.15:
                    print(f"    {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f}")


if __name__ == "__main__":
    main()
