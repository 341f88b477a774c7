#!/usr/bin/env python
"""Turn ncu reports (gpurun_out/*.ncu-rep) into the committed summaries under profiles/.

    python profiles/summarize.py gpurun_out/prof_X.ncu-rep [cells_per_launch]

Writes profiles/<name>_summary.json (selected metrics per profiled launch) and prints them.
"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = [
    'Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def main():
    rep = sys.argv[1]
    cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {}
        for k in WANT:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i] + ' ' + units[i]).strip()
        if cells:
            rd = float(r[hdr.index('dram__bytes_read.sum')]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[hdr.index('dram__bytes_read.sum')]]
            wr = float(r[hdr.index('dram__bytes_write.sum')]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[units[hdr.index('dram__bytes_write.sum')]]
            d['dram_bytes_per_cell'] = (rd + wr) / cells
            d['thread_instructions_per_cell'] = float(r[hdr.index('smsp__inst_executed.sum')]) * 32 / cells
        out.append(d)
    name = os.path.splitext(os.path.basename(rep))[0]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + '_summary.json')
    json.dump(out, open(path, 'w'), indent=1)
    for d in out:
        for k, v in d.items():
            print('%-78s %s' % (k, v))
        print()


if __name__ == '__main__':
    main()
