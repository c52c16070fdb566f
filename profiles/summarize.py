"""Turns the ncu artefacts a gpurun call brings back (gpurun_out/) into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r1_simt.csv > profiles/r1_simt_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_gcl_simt.ncu-rep > profiles/r1_simt_edge_gcl_ncu.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg, tot, n_all = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        if row['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(row['Metric Unit'], v)
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        a = agg.setdefault(name, [0, 0.0, row['Grid Size'], row['Block Size']])
        a[0] += 1; a[1] += v; tot += v; n_all += 1
    print(f'# ncu launch list (gpu__time_duration.sum, --clock-control none): {n_all} launches, {tot / 1e3:.3f} ms total')
    print(f'{"kernel":44s} {"n":>4s} {"total_us":>10s} {"share":>7s} {"avg_us":>9s}  grid / block')
    for k, (n, v, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{k:44s} {n:4d} {v:10.1f} {100 * v / tot:6.1f}% {v / n:9.1f}  {g} / {b}')


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print('== ' + d[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'  {k:72s} {d[i]:>18s} {units[i]}')
        stalls = [(float(d[i]), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and d[i]]
        if stalls:
            print('  top stall reasons (warps stalled per issue-active cycle):')
            for v, h in sorted(stalls, reverse=True)[:6]:
                print(f'    {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""):40s} {v:8.3f}')


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
