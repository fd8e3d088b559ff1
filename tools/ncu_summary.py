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
def summarise(paths):
  for path in paths:
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
  


def opcode_histogram(path, top=16):
    """dynamic SASS opcode mix of the first kernel of a report captured with --import-source / -lineinfo (ncu --page source)"""
    import collections, re
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for n, r in enumerate(rows[:4]):
        if 'Source' in r and 'Instructions Executed' in r:
            hdr = r; rows = rows[n + 1:]; break
    if hdr is None:
        return
    ia, ie = hdr.index('Source'), hdr.index('Instructions Executed')
    ops, tot = collections.Counter(), 0
    for r in rows:
        if len(r) <= ie:
            continue
        try:
            n = int(r[ie] or 0)
        except ValueError:
            continue
        tot += n
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[ia])
        if m:
            ops[m.group(2)] += n
    if tot:
        print("   warp instructions executed: %d ; opcode mix: %s" % (tot, ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(top))))


def sass_counts(path):
    """per-instruction executed counts of the first kernel (index, opcode text, warp instructions executed): joined with nvdisasm -g line
    info by tools/sass_by_line.py"""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for n, r in enumerate(rows[:4]):
        if 'Source' in r and 'Instructions Executed' in r:
            hdr = r; rows = rows[n + 1:]; break
    if hdr is None:
        return
    ia, ie = hdr.index('Source'), hdr.index('Instructions Executed')
    w = csv.writer(sys.stdout)
    for n, r in enumerate(rows):
        if len(r) > ie:
            w.writerow([n, r[ia].strip(), r[ie]])


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    if '--sass-counts' in sys.argv:
        sass_counts(args[0]); sys.exit(0)
    summarise(args)
    if '--opcodes' in sys.argv:
        for a in args:
            opcode_histogram(a)
