R=${1:-r1m}
timeout 300 python -m pytest tests/test_replay_ring.py tests/test_maddpg.py -m gpu -x -q > gpurun_out/${R}_trainer_tests.log 2>&1; tail -5 gpurun_out/${R}_trainer_tests.log
timeout 200 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 1500 --report 300 2>&1 | grep env_steps > gpurun_out/${R}_train_1gpu.log; tail -2 gpurun_out/${R}_train_1gpu.log | cut -c1-120
