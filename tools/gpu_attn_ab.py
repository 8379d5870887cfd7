"""Interleaved A/B timing of two fwb_attn_set_tuning codes on the v1 attention kernel at the hot-path shapes, plus a bit-equality
check of the outputs.  Run under gpurun:

    python tools/gpu_attn_ab.py 0 2        # MUFU-only default vs the speculative single-pass softmax
    python tools/gpu_attn_ab.py 1002 1003  # MUFU ping-pong off / on

Writes gpurun_out/attn_ab_<A>_<B>.log.  For cycle-exact periods use tools/attn_trace.py run 1 <code>."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
import torch
import fwb200

A, B_ = int(sys.argv[1]), int(sys.argv[2])


def timeit(fn, iters=5, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = open(ROOT / "gpurun_out" / f"attn_ab_{A}_{B_}.log", "w")
fwb200.lib.fwb_attn_set_tuning(201)                       # v1 kernel for every head_dim
for (B, H, Lq, Lk, D) in [(1, 40, 32760, 32760, 128), (1, 12, 32760, 32865, 96), (1, 16, 32865, 32865, 64), (1, 40, 4095, 8190, 128)]:
    q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for L in (Lq, Lk, Lk))
    o = torch.empty_like(q)
    fl = 4.0 * B * H * Lq * Lk * D
    res, outs = {A: [], B_: []}, {}
    for rep in range(4):
        for code in (A, B_):
            fwb200.lib.fwb_attn_set_tuning(code)
            res[code].append(timeit(lambda: fwb200.attention(q, k, v, out=o)))
            outs[code] = o.clone()
    line = (f"v1 B{B} H{H} Lq{Lq} Lk{Lk} D{D}: " + "  ".join(f"code {c}: " + "/".join(f"{fl / ms / 1e9:.0f}" for ms in res[c]) + " TF"
                                                             for c in (A, B_)) + f"  outputs equal: {torch.equal(outs[A], outs[B_])}")
    print(line, flush=True)
    out.write(line + "\n")
out.close()
