#!/bin/bash
# GPU call E (8 GPUs): bench at N=8 with CFG-parallel x SP-4 (default) and with SP-8 for the A/B; then N=4
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() {  # n parallel tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $1 --steps 4 --warmup 3 --no-cpu-baseline --parallel $2 > gpurun_out/r02_bench_n$1_$2.json 2> gpurun_out/r02_bench_n$1_$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_n$1_$2.json"))
    print("N=$1 $2:", d["value"], "steps/s", d["ms_per_step"], "ms/step e2e", d["e2e"]["value"], "roofline", d["roofline"] and d["roofline"]["frac"], d["config"]["parallelism"][:60])
except Exception as e:
    print("N=$1 $2 FAILED", e); print(open("gpurun_out/r02_bench_n$1_$2.err").read()[-1500:])
PY
}
run 8 cfg
run 8 sp
run 4 cfg
