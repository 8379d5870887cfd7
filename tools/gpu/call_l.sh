#!/bin/bash
# GPU call L (1 GPU): attention correctness (all configurations) + A/Bs at the hot shapes
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention or bringup" 2>&1 | tail -8)
timeout 300 python tools/gpu_attn_sweep.py 2>&1 | grep -E "multicast|short-kv|^==|PV N" | head -40
