"""Turn an ncu per-launch metrics CSV of one train step (scratch/one_step.py) into a table: one row per launch.
    python profiles/step_table.py gpurun_out/X.csv [out.txt]
When the capture holds several steps only the last one is tabulated."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    by = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = int(r['ID'])
        d = by.setdefault(k, {'name': re.sub(r'\(.*', '', r['Kernel Name']).replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('void ', ''),
                              'grid': r['Grid Size'], 'block': r['Block Size']})
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        m = r['Metric Name']
        if m.startswith('gpu__time'):
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1.0)
            d['us'] = v
        elif 'bytes' in m:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            d[{'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr'}.get(m, 'xbar')] = v
        else:
            d['tensor'] = v
    return by


def last_step(by):
    """scratch/one_step.py runs the step twice (warm-up + the one to read): keep the launches from the last weight re-layout on."""
    ids = [k for k, d in by.items() if d['name'].startswith('k_prep_conv_w')]
    if len(ids) < 2:
        return by
    return collections.OrderedDict((k, d) for k, d in by.items() if k >= ids[-1])


def main():
    by = last_step(load(sys.argv[1]))
    out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
    out.write('# source: %s (ncu, one train step of batch 128 5raw1of; serialised, cold caches: compare shares)\n' % sys.argv[1])
    out.write('%4s %-34s %-12s %9s %9s %9s %9s %7s\n' % ('id', 'kernel', 'grid', 'us', 'dramRdMB', 'dramWrMB', 'xbarMB', 'tensor%'))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    tot = 0.0
    for k, d in by.items():
        out.write('%4d %-34s %-12s %9.1f %9.1f %9.1f %9.1f %7.1f\n' % (k, d['name'][:34], d['grid'].replace(' ', ''), d.get('us', 0), d.get('rd', 0) / 1e6,
                                                                    d.get('wr', 0) / 1e6, d.get('xbar', 0) / 1e6, d.get('tensor', 0)))
        a = agg[d['name']]
        a[0] += 1; a[1] += d.get('us', 0); a[2] += d.get('rd', 0); a[3] += d.get('wr', 0)
        tot += d.get('us', 0)
    out.write('\n# per kernel: launches, total us, share, DRAM read MB, DRAM write MB   (total %.1f us)\n' % tot)
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write('%-34s %4d %9.1f %6.1f%% %9.1f %9.1f\n' % (n[:34], a[0], a[1], 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6))


if __name__ == '__main__':
    main()
