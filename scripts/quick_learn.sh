timeout 300 python -m pytest tests/test_replay_ring.py tests/test_maddpg.py -m gpu -x -q > gpurun_out/r1k_trainer_tests.log 2>&1; tail -5 gpurun_out/r1k_trainer_tests.log
timeout 200 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 1500 --report 300 2>&1 | grep env_steps > gpurun_out/r1k_train_1gpu.log; tail -2 gpurun_out/r1k_train_1gpu.log
timeout 100 python scripts/bench_sampler.py > gpurun_out/r1k_sampler.log 2>&1; cat gpurun_out/r1k_sampler.log
