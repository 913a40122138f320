#!/usr/bin/env python
"""Condense an ncu report (read here, without a GPU) into the JSON summaries kept under profiles/:

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02/flow_cond_ncu.json [extra metric regex ...]

One entry per captured launch with the metrics the roofline arithmetic and the bound analysis use."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    r'^gpu__time_duration\.sum$', r'^dram__bytes_(read|write)\.sum$', r'^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$',
    r'^sm__pipe_tensor.*cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$', r'^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$',
    r'^lts__t_sector_hit_rate\.pct$', r'^launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|cluster_dim_x)$',
    r'^sm__cycles_elapsed\.max$', r'^sm__cycles_active\.avg$', r'^smsp__issue_active\.avg\.pct_of_peak_sustained_active$',
    r'^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$', r'^l1tex__throughput\.avg\.pct_of_peak_sustained_(active|elapsed)$',
    r'^l1tex__data_pipe_lsu_wavefronts(_mem_shared.*)?\.(sum|avg).*$', r'^l1tex__data_bank_.*$',
    r'^sm__inst_executed_pipe_(tensor|uniform|lsu|alu|fma|xu|tmem|uniform_.*|tc.*).*$', r'^smsp__inst_executed_pipe_.*\.sum$',
    r'^smsp__average_warp.*_issue_stalled_.*_per_warp_active\.pct$', r'^smsp__average_warps_issue_stalled_.*$',
    r'^smsp__warp_issue_stalled_.*$', r'^sm__mio.*$', r'^smsp__inst_executed\.sum$', r'^sm__warps_active\.avg\.pct_of_peak_sustained_active$',
    r'^l1tex__m_xbar2l1tex_read_bytes\.sum.*$', r'^lts__t_bytes\.sum.*$', r'^sm__sass_inst_executed_op_shared.*$',
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pats = [re.compile(p) for p in KEEP + sys.argv[3:]]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    names, units = rows[hdr], rows[hdr + 1]
    launches = []
    for r in rows[hdr + 2:]:
        if not r or not r[0].isdigit():
            continue
        d = {'kernel': r[names.index('Kernel Name')]}
        for n, u, v in zip(names, units, r):
            if any(p.search(n) for p in pats):
                try:
                    val = float(v.replace(',', ''))
                except ValueError:
                    continue
                d[n] = {'value': val, 'unit': u}
        launches.append(d)
    json.dump({'source': rep, 'launches': launches}, open(out, 'w'), indent=1)
    for d in launches:
        t = d.get('gpu__time_duration.sum', {})
        print(d['kernel'][:60], t.get('value'), t.get('unit'), len(d) - 1, 'metrics')


if __name__ == '__main__':
    main()
