#!/bin/bash
# GPU call D (2 GPUs): sequence-parallel + CFG-parallel equivalence test (log committed under profiles/), bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
(timeout 900 python -m pytest tests/test_gpu_sp.py -q -x -rA 2>&1 | tail -15) | tee gpurun_out/r02_test_gpu_sp_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n2_cfg.json 2> gpurun_out/r02_bench_n2_cfg.err
head -c 3000 gpurun_out/r02_bench_n2_cfg.json; tail -3 gpurun_out/r02_bench_n2_cfg.err
