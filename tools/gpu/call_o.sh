#!/bin/bash
# GPU call O (1 GPU): ncu --set full of the CURRENT default DiT self-attention launch (K/V multicast pairs, attn_fwd_mc_kernel) for the
# roofline `traffic` record, and of the two row kernels (what bounds them below the copy bandwidth)
mkdir -p gpurun_out
ext() {
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2] if len(r)>2 else r[1]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg.per_second','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__m_xbar2l1tex_read_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__cycles_active.avg','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes.sum.per_second']
for i,n in enumerate(h):
    if any(w == n or (w in n and len(w) > 24) for w in want) or ('tensor' in n and 'pct' in n): print(n, '=', v[i])
"
}
for kind in attn128 ln rms; do
  case $kind in attn128) rx=attn_fwd;; ln) rx=ln_modulate;; rms) rx=rmsnorm_rope;; esac
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/r02o_$kind -f python tools/ncu_attn.py $kind > gpurun_out/r02o_$kind.log 2>&1
  echo "=== $kind"; ext gpurun_out/r02o_$kind.ncu-rep | tee gpurun_out/r02o_$kind.metrics.txt
done
ls -la gpurun_out/r02o_*.ncu-rep
