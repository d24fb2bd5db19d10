#!/usr/bin/env python
"""Print the per-kernel / per-layer table of a bench.py --profile-layers JSON (CUDA-event timings, eager pass)."""
import json
import sys
from collections import defaultdict

d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f'step (sum of launches) {d["step_ms"]:.3f} ms')
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for p in d['launches']:
    a = agg[(p['kernel'], p.get('tag', ''))]
    a[0] += p['ms']
    a[1] += p['flops']
    a[2] += p['bytes']
    a[3] += 1
print(f'{"kernel":30s} {"layer":28s} {"n":>4s} {"ms":>8s} {"share":>6s} {"TFLOP/s":>8s} {"GB/s":>7s}')
for (k, t), (ms, fl, by, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{k:30s} {t:28s} {n:4d} {ms:8.3f} {100 * ms / d["step_ms"]:5.1f}% {fl / ms / 1e9 if ms else 0:8.1f} '
          f'{by / ms / 1e6 if ms else 0:7.0f}')
