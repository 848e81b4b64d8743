#!/usr/bin/env python
"""Extract the per-kernel columns the profiles/ summaries use from an ncu report (`--set full`).

  python tools/ncu_summary.py REPORT.ncu-rep > profiles/rNN_x_ncu_full_summary.csv
  python tools/ncu_summary.py --traffic COMMIT N_CONFIGS REPORT.ncu-rep > profiles/r02_traffic.json
      DRAM bytes (read + write) of ONE pass (the report must hold exactly one pass: `--launch-skip` to the first
      k_broadphase of a step, `-c` = launches per pass), per kernel and in total: what bench.py reports as
      `roofline.traffic`.
"""
import csv
import io
import subprocess
import sys

COLS = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def traffic(commit, n_configs, report):
    import json
    out = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ik, ir, iw, it = (h.index(c) for c in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'))
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}

    def val(r, i):
        return float(r[i].replace(',', '')) * scale.get(units[i], 1.0)
    per, total, narrow = [], 0.0, 0.0
    for r in rows[2:]:
        name = r[ik].split('(')[0].split('<')[0].replace('jg::', '').replace('void ', '').strip()
        b = val(r, ir) + val(r, iw)
        per.append({'kernel': name, 'read': val(r, ir), 'write': val(r, iw), 'time': r[it] + ' ' + units[it]})
        total += b
        if name in ('k_narrowphase', 'k_penetration', 'k_penetration_big', 'k_select'):
            narrow += b
    print(json.dumps({'commit': commit, 'configs_per_pass': int(n_configs), 'whole_pass_bytes': total, 'narrow_phase_bytes': narrow,
                      'bytes_per_check': total / int(n_configs), 'per_kernel_bytes': per,
                      'source': 'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one pass of '
                                '`python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-count --no-extras`'}, indent=1))


def main():
    if sys.argv[1] == '--traffic':
        return traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    idx = [(c, h.index(c)) for c in COLS if c in h]
    w = csv.writer(sys.stdout)
    w.writerow([c for c, _ in idx])
    w.writerow([units[i] for _, i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for _, i in idx])


if __name__ == '__main__':
    main()
