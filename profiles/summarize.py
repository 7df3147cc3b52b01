"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv  profiles/rNN_launches_X.txt  [steps]
    python profiles/summarize.py full     gpurun_out/prof_X.ncu-rep  profiles/rNN_full_X.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']


def launches(src, dst, steps):
    lines = [l for l in open(src) if not l.startswith('==')]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('<unnamed>::', '').replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(row['Metric Unit'], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(dst, 'w') as f:
        f.write('# source: %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)\n' % src)
        f.write('# %d timed steps captured; total %.1f us = %.1f us/step\n' % (steps, tot, tot / steps))
        f.write('%-44s %8s %12s %7s\n' % ('kernel', 'launches', 'total_us', 'share'))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-44s %8d %12.1f %6.1f%%\n' % (k[:44], n, t, 100 * t / tot))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('# source: %s (ncu --set full --clock-control none --import-source on)\n' % src)
        for row in rows[2:]:
            f.write('---- launch id %s\n' % row[0])
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write('%-86s %s %s\n' % (k, row[i][:70], units[i]))
    print(open(dst).read())


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 1)
    else:
        full(sys.argv[2], sys.argv[3])
