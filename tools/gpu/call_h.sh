#!/bin/bash
# GPU call H (1 GPU): ncu evidence for round 2 — launch list of one bench step (shares), --set full capture of the dominant kernel
# (DiT self-attention, attn_fwd_kernel<128,2>) and of the CTA-pair GEMM; heads / VAE gpu tests that changed since call G
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_parity.py -q -x -k "heads or index_exact or vae or joint" 2>&1 | tail -6)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --gpu-reference off --profiler-range > gpurun_out/r02_ncu_bench.json 2> gpurun_out/r02_ncu_bench.err
python - <<'PY'
import csv, collections, json, re
rows = []
with open("gpurun_out/r02_launches.csv", newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for row in r:
    try:
        ns = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row.get("Metric Unit", "ns")
    ms = ns / 1e6 if unit in ("ns", "nsecond") else (ns / 1e3 if unit in ("us", "usecond") else ns)
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:120]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
out = {"note": "ncu --metrics gpu__time_duration.sum --clock-control none over one timed region of bench.py (--profile-from-start off + cudaProfilerStart/Stop around the timed step: launches are serialised and cold-cache, compare SHARES not absolute times)",
       "total_ms": tot, "kernels": [{"name": k, "launches": v[0], "ms": v[1], "share": v[1] / tot} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open("gpurun_out/r02_launch_list_summary.json", "w"), indent=1)
for k in out["kernels"][:14]:
    print(f"{k['share']:.3f} {k['ms']:9.1f} ms x{k['launches']:5d} {k['name'][:90]}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/r02_attn_d128 -f python tools/ncu_attn.py attn128 > /dev/null 2>&1
ncu -i gpurun_out/r02_attn_d128.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2] if len(r)>2 else r[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__cycles_elapsed.avg.per_second','lts__t_bytes.sum','sm__inst_executed_pipe_xu.sum','smsp__inst_executed.sum','l1tex__m_xbar2l1tex_read_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i,n in enumerate(h):
    if any(w in n for w in want) or 'tensor' in n and 'pct' in n: print(n, '=', v[i])
" | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2 -s 1 -c 1 -o gpurun_out/r02_gemm2 -f python tools/ncu_attn.py gemm > /dev/null 2>&1
ncu -i gpurun_out/r02_gemm2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2] if len(r)>2 else r[1]
for i,n in enumerate(h):
    if any(w in n for w in ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__cycles_elapsed.avg.per_second']) or ('tensor' in n and 'pct' in n): print(n, '=', v[i])
" | head -20
ls -la gpurun_out/*.ncu-rep
# reduced-size smoke of the wan22 workload (two experts, expert switch inside the window) before the 8-GPU run
timeout 600 python bench.py --workload wan22_720p --frames 5 --h 16 --w 24 --pcb 2 --irg 2 --steps 2 --warmup 1 --no-cpu-baseline --gpu-reference off > gpurun_out/r02_bench_wan22_reduced.json 2> gpurun_out/r02_bench_wan22_reduced.err; head -c 1200 gpurun_out/r02_bench_wan22_reduced.json; tail -4 gpurun_out/r02_bench_wan22_reduced.err
