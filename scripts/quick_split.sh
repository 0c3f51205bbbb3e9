#!/bin/bash
# dev: the split rollout kernel (FeAR on helper warps): parity tests, then timing with and without it
timeout 300 python -m pytest tests/test_rollout_kernel.py -m gpu -x -q --timeout 200 2>&1 | tail -6
echo "--- split (default)"; timeout 200 python scripts/bench_rollout_kernel.py 2>&1 | grep "E=4096\|E=16384" | head -7
echo "--- GW_ROLL_SPLIT=0"; GW_ROLL_SPLIT=0 timeout 200 python scripts/bench_rollout_kernel.py 2>&1 | grep "E=4096" | head -4
