#!/bin/bash
# GPU call Q (1 GPU): 256-thread row CTAs — tests and A/B timings
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "row_kernels or ln_modulate or rmsnorm" 2>&1 | tail -5)
timeout 300 python tools/gpu_row_ab.py 2>&1 | tee gpurun_out/r02_row_ab_threads.log
