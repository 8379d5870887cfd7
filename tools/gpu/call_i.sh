#!/bin/bash
# GPU call I (8 GPUs): BASELINE configs[3] — Wan2.2-Fun-A14B-Control-Camera shape, 720p latents, two experts resident, CFG-parallel x SP-4
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload wan22_720p --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_wan22_720p_n8.json 2> gpurun_out/r02_bench_wan22_720p_n8.err
head -c 3000 gpurun_out/r02_bench_wan22_720p_n8.json; tail -5 gpurun_out/r02_bench_wan22_720p_n8.err | cut -c1-300
nvidia-smi --query-gpu=memory.used --format=csv | head -3
