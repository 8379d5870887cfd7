#!/bin/bash
# GPU call B: attention correctness with the packed softmax, variant sweep, full reference parity, head index tests
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | tail -8)
timeout 300 python tools/gpu_attn_sweep.py 2>&1 | tail -100
(timeout 900 python -m pytest tests/test_gpu_ref_parity.py -q -s 2>&1 | grep -E "ref-parity|passed|failed|Error|error" | cut -c1-600) > gpurun_out/r02_ref_parity.log 2>&1
cat gpurun_out/r02_ref_parity.log
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "index_exact or with_heads" 2>&1 | tail -5)
