#!/bin/bash
# GPU call G (1 GPU): full gpu test suite on the final kernels + headline bench (with gpu_reference and CPU baseline)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) | tee gpurun_out/r02_pytest_gpu.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; head -c 3500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
