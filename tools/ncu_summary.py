#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full` report: duration, DRAM traffic, throughput and occupancy metrics.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem/block"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("launch__waves_per_multiprocessor", "waves"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    print(f"{name}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
    parts = []
    for key, label in want:
        if key in idx:
            parts.append(f"{label}={r[idx[key]]} {units[idx[key]]}".strip())
    print("   " + "; ".join(parts))
