# batch sizes as multiples of the SM count: usage bash tools/sweep_batch.sh
for cfg in "128 640" "148 740" "148 2960" "296 740"; do
  set -- $cfg
  YNET_MAX_STACKED_PASSES=$2 timeout 600 python bench.py --agents $1 --steps 12 --warmup 3 --no-cpu-baseline --no-roofline --torch-cuda-agents 0 > /tmp/sweep.log 2>&1
  python - "$1" "$2" <<'P'
import json, sys
line = [l for l in open('/tmp/sweep.log') if l.startswith('{')]
if not line:
    print('agents', sys.argv[1], 'passes', sys.argv[2], 'FAILED', open('/tmp/sweep.log').read()[-300:])
else:
    d = json.loads(line[-1])
    print('agents', sys.argv[1], 'passes', sys.argv[2], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'GB', round(d['diag']['mem']['allocated_peak_gb'], 1))
P
done
