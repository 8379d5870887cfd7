#!/bin/bash
# GPU call N (1 GPU): in-step A/Bs of the attention knobs under the power cap (one process each, alternating 3-step segments)
mkdir -p gpurun_out; rm -f gpurun_out/r02_step_ab.log
timeout 900 python tools/gpu_step_ab.py fwb_attn_set_exp2_poly 0 2 --rounds 3 --steps 3 2>&1 | tail -1
timeout 900 python tools/gpu_step_ab.py fwb_attn_set_variant 0 2 1 --rounds 2 --steps 3 2>&1 | tail -1
