#!/bin/bash
# GPU call C (1 GPU): attention timeline traces (poly 0 / 2 / 3), GEMM correctness after the epilogue change, per-kernel step
# breakdown, vggt_only workload, joint parity rerun
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" 2>&1 | tail -6)
timeout 300 python tools/gpu_attn_sweep.py 2>&1 | tail -60
timeout 120 python tools/attn_trace.py run 3 2 2>&1 | tail -16
for P in 0 2 3; do timeout 120 python tools/attn_trace.py run 1 $P 2>&1 | tail -16; done
timeout 120 python tools/attn_trace.py run 2 2 2>&1 | tail -16
(timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "gemm or linear" 2>&1 | tail -5)
timeout 900 python bench.py --steps 2 --warmup 2 --breakdown --no-cpu-baseline --gpu-reference off > gpurun_out/r02_bench_breakdown.json 2> gpurun_out/r02_bench_breakdown.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_breakdown.json'))
print("ms/step", d["ms_per_step"], "roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
for r in d["breakdown_ms_per_step"][:34]:
    print(f"{r['ms_per_step']:8.1f} ms x{r['launches_per_step']:6.1f} {r['tag']}")
PY
timeout 600 python bench.py --workload vggt_only --steps 3 --warmup 2 > gpurun_out/r02_bench_vggt_only.json 2> gpurun_out/r02_bench_vggt_only.err; head -c 2500 gpurun_out/r02_bench_vggt_only.json; tail -3 gpurun_out/r02_bench_vggt_only.err
(timeout 900 python -m pytest tests/test_gpu_ref_parity.py -q -s -k "joint or drift" 2>&1 | grep -E "ref-parity|passed|failed|Error|error" | cut -c1-500)
