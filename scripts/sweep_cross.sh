#!/bin/bash
# dev: small-batch kernel vs thread-per-env kernel (32- and 256-env tiles) across batch sizes
for E in 1024 2048 4096 8192 16384 32768 65536; do
  for mode in "small GW_SMALL=1 GW_TILE=32" "big32 GW_SMALL=0 GW_TILE=32" "big256 GW_SMALL=0 GW_TILE=256"; do set -- $mode
    env $2 $3 python bench.py --envs $E --fear ${FEAR:-1} --steps 640 --warmup 64 --no-cpu-baseline --no-e2e --no-scale-points 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('E %7d %-7s: %8.3f us/step  %.3f G agent-steps/s' % ($E, '$1', d['ms_per_step']*1e3, d['value']/1e9))
"
  done
done
