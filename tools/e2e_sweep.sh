#!/bin/bash
# Developer tool (GPU box): e2e leg of bench.py under different streaming parameters.
cd "$(dirname "$0")/.."
run() {
  echo "== $*"
  timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.2fM  e2e %.2fM  e2e ms %.1f kernel-sum %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['kernel_ms_sum_per_step']))
"
}
run
run --no-lpt --no-ramp
run --no-ramp
run --block-ligands 65536
run --block-ligands 32768
run --block-ligands 65536 --no-ramp
