#!/bin/bash
# GPU call P (1 GPU): bulk-copy row ring — bit-identity test, compute-sanitizer (memcheck + racecheck) on a small ring run, A/B timings
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "ring or ln_modulate or rmsnorm" 2>&1 | tail -5)
cat > /tmp/ring_small.py <<'PY'
import sys; sys.path.insert(0, "fantasy-world_b200")
import torch, fwb200
fwb200.lib.fwb_rowwise_set_ring(1)
x = torch.randn(700, 5120, device="cuda"); xb = x.to(torch.bfloat16)
w = torch.randn(5120, device="cuda"); cs = torch.randn(700, 64, 2, device="cuda")
fwb200.ln_modulate(x, eps=1e-6, mul=w, add=w); fwb200.ln_modulate(xb, eps=1e-6, w=w, b=w)
fwb200.rmsnorm_rope_(xb, w=w, eps=1e-6, cos_sin=cs, head_dim=128); fwb200.rmsnorm_rope_(xb, cos_sin=cs, head_dim=128)
torch.cuda.synchronize(); print("ok")
PY
for tool in memcheck racecheck; do timeout 300 compute-sanitizer --tool $tool --kernel-regex kns=ln_modulate --kernel-regex kns=rmsnorm_rope python /tmp/ring_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|^ok" | head -8; done
timeout 300 python tools/gpu_row_ab.py 2>&1 | tee gpurun_out/r02_row_ab.log
