#!/usr/bin/env python
"""Per-kernel DRAM traffic of one bench step from an ncu CSV log.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --profile-from-start off -c 600 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 ...
    python tools/ncu_traffic.py gpurun_out/traffic.csv <agents per step> > profiles/ncu_traffic_r02.json

bench.py reads the JSON for `roofline.traffic` (average DRAM bytes per launch of the dominant kernel).
"""
import csv
import json
import re
import sys

LABELS = [   # ops.py timing label <- ncu kernel name
    ('tc_rowconv2_kernel<pred,softargmax>', r'tc_rowconv2_kernel<\(bool\)1|tc_rowconv2_kernel<true|tc_rowconv2_kernel<1'),
    ('tc_rowconv2_kernel', r'tc_rowconv2_kernel'),
    ('tc_rowconv_kernel<pred,softargmax>', r'tc_rowconv_kernel<\(bool\)1|tc_rowconv_kernel<true|tc_rowconv_kernel<1'),
    ('tc_rowconv_kernel', r'tc_rowconv_kernel<\(bool\)0|tc_rowconv_kernel<false|tc_rowconv_kernel<0'),
    ('tc_conv3x3_kernel', r'tc_conv_kernel<\d, 9, 0>'),
    ('tc_conv3x3_hilo_kernel', r'tc_conv_kernel<\d, 9, 4>'),
    ('tc_upconv3x3_kernel', r'tc_conv_kernel<\d, 9, 3>|upconv_ringfix'),
    ('tc_pred_softargmax_kernel', r'tc_pred_softargmax_kernel'),
    ('tc_conv_kernel<1x1,f32>', r'tc_conv_kernel<\d, 1, 1>'),
    ('wp_pyramid_c8_kernel', r'wp_pyramid'),
    ('kmeans_kernel', r'kmeans_kernel'),
    ('cdf_sequential_kernel+cdf_search_kernel', r'cdf_sequential_kernel|cdf_search_kernel'),
    ('cws_partial_kernel', r'cws_partial_kernel'),
]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3,
        'nsecond': 1e-6, 'msecond': 1.0, 'second': 1e3}


def main(path, agents):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    idx = {k: j for j, k in enumerate(rows[h])}
    per = {}
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]):
            continue
        name = r[idx['Kernel Name']].split('(')[0].replace('(int)', '')
        lab = next((l for l, rx in LABELS if re.search(rx, name)), name.replace('void ynet::', '').replace('ynet::', ''))
        d = per.setdefault(lab, {'ids': set(), 'dram_bytes': 0.0, 'ms': 0.0})
        v = float(r[idx['Metric Value']].replace(',', '')) * UNIT.get(r[idx['Metric Unit']], 1.0)
        m = r[idx['Metric Name']]
        if m.startswith('dram__bytes'):
            d['dram_bytes'] += v
        elif m == 'gpu__time_duration.sum':
            d['ms'] += v
            d['ids'].add(r[idx['ID']])
    out = {'source': path, 'agents_per_step': agents, 'note': 'one graph-replayed bench step under ncu (cold-cache, '
           'serialised launches); dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum summed over the launches',
           'kernels': {k: {'launches': len(v['ids']), 'dram_bytes': v['dram_bytes'], 'ms': v['ms'],
                           'gb_per_s': v['dram_bytes'] / v['ms'] / 1e6 if v['ms'] else None}
                       for k, v in sorted(per.items(), key=lambda kv: -kv[1]['ms'])}}
    json.dump(out, sys.stdout, indent=1)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]))
