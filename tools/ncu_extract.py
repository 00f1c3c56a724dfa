#!/usr/bin/env python
"""Summarises an ``ncu --set full --import-source on`` report for profiles/

    python tools/ncu_extract.py REPORT.ncu-rep KERNEL_REGEX SOURCE.cu [--so LIB] > profiles/xxx.txt

Prints (1) the launch/occupancy/throughput/stall metrics of the raw page for
every captured launch whose name matches, (2) the executed warp instructions
and stall samples aggregated per source line: the SASS listing of the built
library (``cuobjdump -xelf`` + ``nvdisasm -g -c``, needs -lineinfo) is zipped
with the per-instruction rows of ``ncu --page source --csv``.  Analysis runs
on the container (no GPU needed).
"""

import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

RAW_KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__occupancy_limit_warps',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
]


def ncu_page(report, page):
    out = subprocess.run(['ncu', '-i', report, '--page', page, '--csv'],
                         capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def sass_lines(lib, cu_name, kernel_regex):
    """[(file, line)] of every SASS instruction of the first matching kernel"""
    tmp = tempfile.mkdtemp(prefix='ncu_extract_')
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)],
                   cwd=tmp, capture_output=True, check=True)
    stem = os.path.splitext(os.path.basename(cu_name))[0]
    cubin = [f for f in os.listdir(tmp) if f.startswith(stem + '.')][0]
    text = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)],
                          capture_output=True, text=True, check=True).stdout
    lines = text.split('\n')
    start = [i for i, l in enumerate(lines)
             if '.section' in l and '.text.' in l and re.search(kernel_regex, l)]
    if not start:
        raise SystemExit('kernel not found in the SASS listing')
    seq, cur = [], None
    for l in lines[start[0] + 1:]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l):
            seq.append(cur)
        if '.section' in l and '.text.' in l and len(seq) > 16:
            break
    return seq


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('kernel')
    ap.add_argument('source')
    ap.add_argument('--so', default='shennong_b200/_build/libsnb.so')
    ap.add_argument('--top', type=int, default=30)
    ap.add_argument('--sass', default=None,
                    help='regex on the mangled name (template instances)')
    ap.add_argument('--column', default=None,
                    help='also rank the source lines by this source-page '
                         'column, e.g. "L1 Wavefronts Shared"')
    a = ap.parse_args()

    raw = ncu_page(a.report, 'raw')
    head, units = raw[0], raw[1]
    name_col = head.index('Kernel Name')
    print(f'# {os.path.basename(a.report)} -- raw page')
    for row in raw[2:]:
        if not re.search(a.kernel, row[name_col]):
            continue
        print(f'## {row[name_col]}')
        for key in RAW_KEYS:
            if key in head:
                i = head.index(key)
                print(f'{key},{row[i]},{units[i]}')

    src = ncu_page(a.report, 'source')
    hdr = [i for i, r in enumerate(src) if 'Instructions Executed' in r][0]
    cols, data = src[hdr], src[hdr + 1:]
    ie, isamp = cols.index('Instructions Executed'), cols.index('# Samples')
    seq = sass_lines(a.so, a.source, a.sass or a.kernel)
    print(f'\n# source page: {len(data)} SASS instructions in the report, '
          f'{len(seq)} in the listing of {a.so}')
    if len(seq) != len(data):
        print('# WARNING: the library was rebuilt since the capture; '
              'per-line attribution skipped')
        return
    inst, samp = collections.Counter(), collections.Counter()
    for loc, r in zip(seq, data):
        inst[loc] += int(r[ie])
        samp[loc] += int(r[isamp])
    ti, ts = sum(inst.values()), sum(samp.values())
    text = open(a.source).read().split('\n')
    print(f'# warp instructions executed: {ti}; stall samples: {ts}')
    print('inst_pct,stall_pct,file:line,source')
    for loc, _ in sorted(samp.items(), key=lambda kv: -kv[1])[:a.top]:
        line = ''
        if loc and loc[0] == os.path.basename(a.source):
            line = text[loc[1] - 1].strip()[:90]
        where = f'{loc[0]}:{loc[1]}' if loc else '?'
        print(f'{100 * inst[loc] / ti:.1f},{100 * samp[loc] / ts:.1f},'
              f'{where},"{line}"')
    if a.column:
        ic = cols.index(a.column)
        extra = collections.Counter()
        for loc, r in zip(seq, data):
            try:
                extra[loc] += int(r[ic])
            except ValueError:
                pass
        te = sum(extra.values())
        print(f'\n# {a.column}: {te} in total')
        print('pct,value,file:line,source')
        for loc, v in sorted(extra.items(), key=lambda kv: -kv[1])[:a.top]:
            line = ''
            if loc and loc[0] == os.path.basename(a.source):
                line = text[loc[1] - 1].strip()[:90]
            where = f'{loc[0]}:{loc[1]}' if loc else '?'
            print(f'{100 * v / max(te, 1):.1f},{v},{where},"{line}"')


if __name__ == '__main__':
    main()
