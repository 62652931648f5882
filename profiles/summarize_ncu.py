"""Summarises ncu output into the small text files committed under profiles/.
  python profiles/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
  python profiles/summarize_ncu.py full gpurun_out/prof_gram.ncu-rep > profiles/rNN_gram_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
    'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg.per_second',
    'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi, bi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Block Size')
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        agg.setdefault((r[ki][:100], r[gi], r[bi]), []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)')
    print(f'{"kernel":102s} {"grid":>16s} {"block":>14s} {"n":>3s} {"avg_us":>12s} {"share":>7s}')
    for (k, g, b), v in agg.items():
        print(f'{k:102s} {g:>16s} {b:>14s} {len(v):3d} {sum(v) / len(v) / 1e3:12.1f} {sum(v) / tot:7.3f}')


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print('# ncu --set full --clock-control none; one column per captured launch')
    for want in ['Kernel Name'] + KEYS:
        for i, h in enumerate(hdr):
            if h == want:
                print(f'{h:88s} {units[i]:14s} ' + '  '.join(r[i] for r in data))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
