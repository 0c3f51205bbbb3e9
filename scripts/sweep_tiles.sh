#!/bin/bash
# dev: thread-per-env kernel with 32/64/128/256-env tiles across batch sizes
for E in 8192 16384 32768 65536 131072 262144; do
  for T in 32 64 128 256; do
    GW_SMALL=0 GW_TILE=$T python bench.py --envs $E --fear ${FEAR:-1} --steps 640 --warmup 64 --no-cpu-baseline --no-e2e --no-scale-points 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('E %7d tile %3d: %8.3f us/step  %.3f G agent-steps/s' % ($E, $T, d['ms_per_step']*1e3, d['value']/1e9))
"
  done
done
