#!/bin/bash
# GPU call K (2 GPUs): row-kernel sweep after the occupancy-based grid, then the driver's own invocations at N=2: our arm with all
# legs (CPU baseline after the collective teardown) and the reference arm under torchrun
mkdir -p gpurun_out
bash tools/gpu/call_j.sh 2>&1 | tail -9
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_n2_driverlike.json 2> gpurun_out/r02_bench_n2_driverlike.err; echo "ours rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2_driverlike.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline'], d['gpu_reference'])" | cut -c1-600
tail -3 gpurun_out/r02_bench_n2_driverlike.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-budget 10 --no-cpu-full > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; echo "ref rc=$?"; head -c 600 gpurun_out/r02_bench_ref_n2.json
