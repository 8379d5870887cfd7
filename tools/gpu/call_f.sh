#!/bin/bash
# GPU call F (1 GPU): attention / bring-up correctness (incl. native head_dim-96 PV width) + variant sweep
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_ops.py -x -q -k "attention or bringup" 2>&1 | tail -12)
timeout 300 python tools/gpu_attn_sweep.py --quick 2>&1 | tail -40
