#!/bin/bash
# One short gpurun call: the replay-sampler tests, the trainer tests, training-loop throughput with and without the
# fused sampler, and bench lines of BASELINE configs 3 and 4 (parity-test cases; measured here for the record).
R=${1:-r1j}
timeout 300 python -m pytest tests/test_replay_ring.py tests/test_maddpg.py -m gpu -x -q > gpurun_out/${R}_sampler_tests.log 2>&1; tail -5 gpurun_out/${R}_sampler_tests.log
timeout 200 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 1200 --report 300 2>&1 | grep env_steps > gpurun_out/${R}_train_1gpu.log; tail -1 gpurun_out/${R}_train_1gpu.log
timeout 200 python -m marl_responsible_nav_b200.train --torch-sampler --config custom_fear_10 --envs 4096 --steps 1200 --report 300 2>&1 | grep env_steps > gpurun_out/${R}_train_1gpu_torch_sampler.log; tail -1 gpurun_out/${R}_train_1gpu_torch_sampler.log
timeout 100 python scripts/bench_sampler.py > gpurun_out/${R}_sampler.log 2>&1; cat gpurun_out/${R}_sampler.log
X="--no-cpu-baseline --no-e2e --no-scale-points --steps 1000 --warmup 100"
python bench.py --envs 16384 --obs bf16 $X > gpurun_out/${R}_bench_config3_cnn16384_bf16.json 2>/dev/null
python bench.py --envs 16384 --obs f32 $X > gpurun_out/${R}_bench_config3_cnn16384_f32.json 2>/dev/null
python bench.py --envs 1048576 --scenario "Level 5" $X > gpurun_out/${R}_bench_config4_level5_1m.json 2>/dev/null
python bench.py --envs 1048576 --scenario "GameMap" $X > gpurun_out/${R}_bench_config4_gamemap_1m.json 2>/dev/null
python bench.py --envs 125000 $X > gpurun_out/${R}_bench_config4_125000.json 2>/dev/null
for f in gpurun_out/${R}_bench_config*.json; do python -c "
import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel'))"; done
