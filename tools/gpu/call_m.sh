#!/bin/bash
# GPU call M (1 GPU): in-step A/B of the multicast pairs (one process, alternating segments), then of the exp2 polynomial share
mkdir -p gpurun_out; rm -f gpurun_out/r02_step_ab.log
(timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "multicast" 2>&1 | tail -3)
timeout 900 python tools/gpu_step_ab.py fwb_attn_set_multicast 0 1 --rounds 3 --steps 3 2>&1 | tail -2
