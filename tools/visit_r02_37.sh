#!/bin/bash
# A/B on ONE box: fine-tune bench at HEAD and at the commit before the host-side changes (worktree _old, same .so)
mkdir -p gpurun_out
for tree in . _old . _old; do
  for B in fp32 bf16x3; do
    (cd $tree && timeout 600 python bench.py --mode finetune --workload sdd_short --agents 30 --steps 10 --warmup 3 --backend $B 2>/dev/null | grep "^{" | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tree', '$B', round(d['ms_per_step'],2))")
  done
done | tee gpurun_out/finetune_ab.log
