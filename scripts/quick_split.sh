#!/bin/bash
# dev: FeAR on helper warps (the *_split_kernel's): parity tests, then the bench's headline / e2e / per-step-launch figures with and without
timeout 600 python -m pytest tests/test_rollout_kernel.py tests/test_gpu_parity.py tests/test_replay_ring.py -m gpu -x -q --timeout 300 2>&1 | tail -6
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
for v in 1 0; do
  echo "--- GW_SPLIT=$v"
  GW_SPLIT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scale-points --no-train 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); e = d['e2e']
        print('value %.3f G  (%.2f us/step)  per-step-launch %.3f G  e2e %.1f M  launch-per-step e2e %.1f M  memcpy %.1f M' % (d['value']/1e9, d['ms_per_step']*1e3, d['launch_per_step']['value']/1e9, e['value']/1e6, e.get('launch_per_step_value',0)/1e6, e.get('memcpy_value',0)/1e6))
"
done
