#!/bin/bash
# dev: parity tests + short bench points (4096 / 1M envs, FeAR on / off) + phase trace; one gpurun call
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for spec in "4096 1" "1048576 1" "1048576 0" "16384 1" "65536 1"; do set -- $spec
  python bench.py --envs $1 --fear $2 --steps 640 --warmup 64 --no-cpu-baseline --no-e2e --no-scale-points 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('envs %8d fear %d: %.3f us/step  %.3f G agent-steps/s  frac %.3f' % ($1, $2, d['ms_per_step']*1e3, d['value']/1e9, d['roofline']['frac']))
    else: print(l.rstrip())
"
done
python scripts/trace_phases.py 4096 1
