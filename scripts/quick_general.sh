#!/bin/bash
# dev: general-layout parity tests + its bench extra; one gpurun call
timeout 900 python -m pytest tests/test_general_layout.py -m gpu -x -q --timeout 600 2>&1 | tail -15
timeout 300 python - <<'PY' 2>&1 | tail -20
import json, sys
sys.argv = ["bench.py"]
import bench
a = bench.parse()
import torch
r = bench.general_layout_extra(a, torch.device("cuda:0"))
for pt in r["points"]:
    print("envs %6d fear %d: %.1f us/step  %.1f M agent-steps/s  %.0f GB/s  frac %.3f  tasks %.2f" % (
        pt["envs"], pt["fear"], pt["ms_per_step"] * 1e3, pt["agent_steps_per_s"] / 1e6, pt["achieved_gbs"], pt["frac_of_hbm_peak"], pt["fear_tasks_per_env_step"]))
print(r["cpu_baseline"])
PY
python scripts/ncu_general.py 1 65536 > /dev/null 2>&1 && timeout 300 ncu --set full --import-source on --clock-control none -k regex:gww_step -s 8 -c 1 -o gpurun_out/r2p_gww_fear1 -f python scripts/ncu_general.py 1 65536 > gpurun_out/r2p_gww_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gww_step -s 8 -c 1 -o gpurun_out/r2p_gww_fear0 -f python scripts/ncu_general.py 0 65536 >> gpurun_out/r2p_gww_ncu.log 2>&1
tail -3 gpurun_out/r2p_gww_ncu.log
