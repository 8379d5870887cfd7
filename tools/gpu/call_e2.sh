#!/bin/bash
# GPU call E2 (8 GPUs): final-code scaling point, CFG-parallel x SP-4, driver-like flags (with the CPU baseline leg on rank 0)
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n8_final.json')); print('N=8:', d['value'], 'steps/s', d['ms_per_step'], 'ms e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['clocks'], d['cpu_baseline'])" | cut -c1-700
tail -3 gpurun_out/r02_bench_n8_final.err | cut -c1-300
