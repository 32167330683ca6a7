#!/bin/bash
# quick stage timings of one workload for each tile-kernel launch shape: profiles/tools/quick_bench.sh cfg3 "0 1 2"
wl=${1:-cfg3}; shift
for v in ${1:-0}; do
  LCR_TILE_VARIANT=$v python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('variant $v', '$wl', 'ms/step %.3f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], {k:round(x,3) for k,x in d['kernel_ms_per_step'].items()}, 'frac %.3f'%d['roofline']['frac'], 'attempts', d['config']['run_attempts_per_step'])"
done
