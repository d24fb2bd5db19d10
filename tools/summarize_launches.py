#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table.

    python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/launches_rNN_summary.md

ncu serialises launches and runs them cold-cache, so only each kernel's SHARE is comparable with the
CUDA-event numbers of bench.py (B200_PROFILING.md).
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r'\(.*$', '', name)                 # drop the argument list
    name = name.replace('ynet::', '')
    return name


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        ns = float(r['Metric Value'].replace(',', ''))
        if r.get('Metric Unit') == 'us':
            ns *= 1e3
        elif r.get('Metric Unit') == 'ms':
            ns *= 1e6
        rows.append((int(r['ID']), short(r['Kernel Name']), ns))
    rows = [r for r in rows if r[0] >= skip]
    agg = defaultdict(lambda: [0, 0.0])
    for _, k, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f'# ncu launch list summary: {path}')
    print(f'\n{len(rows)} launches (IDs >= {skip}), {total / 1e6:.3f} ms of serialised device time\n')
    print('| kernel | launches | total ms | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |')


if __name__ == '__main__':
    main()
