#!/usr/bin/env python
"""Top warp-stall sites (SASS level) of one launch in an .ncu-rep captured with --import-source on.

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep <launch index> [top N]
"""
import csv
import io
import subprocess
import sys

path, skip = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(skip),
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:100])
hdr = rows[1]
idx = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
def num(r, k):
    try:
        return int(float(r[idx[k]] or 0))
    except ValueError:
        return 0
tot = sum(num(r, '# Samples') for r in data)
print('total samples', tot, ' instructions', len(data))
for r in sorted(data, key=lambda r: -num(r, '# Samples'))[:top_n]:
    n = num(r, '# Samples')
    stalls = {k: num(r, k) for k in hdr if k.startswith('stall_') and '(' not in k}
    s = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"{n:6d} {100 * n / tot:5.1f}% {r[idx['Address']][-5:]} {r[idx['Source']][:64]:64s} {s}")
