#!/bin/bash
# Dev tool: sweep traverse-kernel tuning knobs on the bench workload (run on the GPU box).
for B in 4 6; do for R in 8 16 24; do for L in 1 2 4 8; do
  echo -n "blocks=$B refill=$R leaf=$L: "
  MGODPL_TRAVERSE_BLOCKS=$B MGODPL_REFILL_MIN=$R MGODPL_LEAF_MIN=$L python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  %.3g states/s' % (d['ms_per_step'], d['value']))"
done; done; done
