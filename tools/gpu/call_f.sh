#!/bin/bash
# GPU call F (1 GPU): variant 3 (pipelined softmax) correctness + sweep
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | tail -12)
timeout 300 python tools/gpu_attn_sweep.py 2>&1 | tail -60
timeout 120 python tools/attn_trace.py run 3 2 2>&1 | tail -16
