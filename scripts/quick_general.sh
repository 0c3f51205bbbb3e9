#!/bin/bash
# dev: general-layout parity tests + its bench extra; one gpurun call
timeout 900 python -m pytest tests/test_general_layout.py -m gpu -x -q --timeout 600 2>&1 | tail -15
timeout 300 python - <<'PY' 2>&1 | grep -v '^ *["{}[]]' ; timeout 300 python - <<'PY' 2>&1 | tail -70
import json, sys, types
sys.argv = ["bench.py"]
import bench
a = bench.parse()
import torch
r = bench.general_layout_extra(a, torch.device("cuda:0"))
print(json.dumps(r, indent=1))
PY
