"""Aggregate a bench.py --profile-layers dump: per (kernel, shape) device time, TFLOP/s, GB/s."""
import json
import sys
from collections import defaultdict

d = json.load(open(sys.argv[1]))
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for p in d['launches']:
    a = agg[(p['kernel'], p['tag'])]
    a[0] += p['ms']; a[1] += p['flops']; a[2] += p['bytes']; a[3] += 1
print(f"step_ms (sum of profiled launches) {d['step_ms']:.3f}")
bykern = defaultdict(float)
for (k, t), a in agg.items():
    bykern[k] += a[0]
for k, v in sorted(bykern.items(), key=lambda kv: -kv[1]):
    print(f'  {v:9.3f} ms  {100 * v / d["step_ms"]:5.1f}%  {k}')
print()
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tf = a[1] / a[0] / 1e9 if a[0] else 0
    gb = a[2] / a[0] / 1e6 if a[0] else 0
    print(f'{a[0]:9.3f} ms n={a[3]:3d} {tf:8.1f} TF/s {gb:8.1f} GB/s  {k[0]} {k[1]}')
