#!/bin/bash
# GPU call R (1 GPU): row kernels with one barrier per reduction — row tests + timings
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "ln_modulate or rmsnorm or ln64" 2>&1 | tail -3)
timeout 300 python tools/gpu_row_ab.py 2>&1 | tee gpurun_out/r02_row_ab_1barrier.log
