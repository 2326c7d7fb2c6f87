"""Prints one line per profiled kernel of an .ncu-rep (ncu --set full): time, DRAM bytes, occupancy, pipe utilisation, top stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
out = open(rep, errors="ignore").read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]
want = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'blk'), ('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrMB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'), ('launch__registers_per_thread', 'regs'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu%'), ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma%'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'), ('smsp__inst_executed.sum', 'winst'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'longsb'), ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'shortsb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'bar'), ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'math'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'notsel'), ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'noinst'),
        ('lts__t_sector_hit_rate.pct', 'l2hit%')]
print(' | '.join(s for _, s in want))
for r in rows[2:]:
    vals = []
    for n, s in want:
        v = r[hdr.index(n)] if n in hdr else '?'
        if s == 'kernel': v = v[:40]
        else:
            try: v = '%.1f' % float(v.replace(',', ''))
            except ValueError: pass
        vals.append(v)
    print(' | '.join(vals))
