#!/bin/bash
# two GPUs (gpurun --gpus 2): the full GPU suite incl. the two-rank training test, then bench.py at the driver's flags
R=${1:-r2q}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${R}_gpu_tests_2gpu.log 2>&1; tail -2 gpurun_out/${R}_gpu_tests_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${R}_bench_2gpu.json 2> gpurun_out/${R}_bench_2gpu.err; tail -c 400 gpurun_out/${R}_bench_2gpu.json
