#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep (DRAM bytes, duration, issue, occupancy limits, shared-memory bank
conflicts, stall reasons) in the format of profiles/r01_ncu_*.txt.  usage: ncu_detail.py report.ncu-rep [more...]"""
import csv, subprocess, sys, io
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:150], d.get("Grid Size"), d.get("Block Size")))
        for k in KEEP:
            if k in d: print("   %-92s %s %s" % (k, d[k], u.get(k, "")))
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
                try:
                    if float(d[k]) >= 0.3: print("   %-92s %s" % (k, d[k]))
                except ValueError:
                    pass
