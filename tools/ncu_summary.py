#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into one table: duration, DRAM bytes, throughput %, occupancy, top stalls."""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'local_load_requests', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(path, "empty"); continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", path.split('/')[-1], '|', r[hdr.index('Kernel Name')][:90])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("   %-95s %s %s" % (w, r[i], units[i]))
        break
