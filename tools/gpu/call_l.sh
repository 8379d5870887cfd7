#!/bin/bash
# GPU call L (1 GPU): cluster-multicast attention — bit-identity tests, A/B at the hot shapes
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_ops.py -x -q -k "multicast or kernel_variants" 2>&1 | tail -8)
timeout 300 python tools/gpu_attn_sweep.py 2>&1 | grep -E "multicast|^==|PV N" | head -30
