#!/bin/bash
# GPU call S (1 GPU): final-tree validation — full gpu suite, smoke(), short bench line (no CPU / reference legs)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4) | tee gpurun_out/r02_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --gpu-reference off > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err; head -c 1500 gpurun_out/r02_bench_n1_final2.json; tail -2 gpurun_out/r02_bench_n1_final2.err
