#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (run here, no GPU needed):  python tools/summarise_ncu.py rep.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(d.get("dram__bytes_read.sum", 0) or 0) * scale.get(u.get("dram__bytes_read.sum", "byte"), 1.0) / 1e6
        wr = float(d.get("dram__bytes_write.sum", 0) or 0) * scale.get(u.get("dram__bytes_write.sum", "byte"), 1.0) / 1e6
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k} = {d[k]} {u[k]}")
        print(f"  dram traffic (read+write) = {rd + wr:.3f} Mbyte")
        for k in hdr:
            if "issue_stalled" in k and "per_issue_active" in k:
                try:
                    if float(d[k]) >= 0.1:
                        print(f"  stall {k.split('issue_stalled_')[1].split('_per_issue')[0]} = {float(d[k]):.3f} warps/issue")
                except ValueError:
                    pass
        print()


if __name__ == "__main__":
    main(sys.argv[1])
