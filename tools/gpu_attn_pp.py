"""A/B of the v1 attention kernel with / without the MUFU ping-pong (fwb_attn_set_tuning 1002 / 1003), interleaved runs."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
import torch
import fwb200


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


out = open(ROOT / "gpurun_out" / "attn_pp.log", "w")
fwb200.lib.fwb_attn_set_tuning(201)
for (B, H, Lq, Lk, D) in [(1, 40, 32760, 32760, 128), (1, 12, 32760, 32865, 96), (1, 16, 32865, 32865, 64)]:
    q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for L in (Lq, Lk, Lk))
    o = torch.empty_like(q)
    fl = 4.0 * B * H * Lq * Lk * D
    res = {0: [], 1: []}
    for rep in range(4):
        for pp in (0, 1):
            fwb200.lib.fwb_attn_set_tuning(1002 + pp)
            res[pp].append(timeit(lambda: fwb200.attention(q, k, v, out=o)))
    line = f"v1 B{B} H{H} Lq{Lq} Lk{Lk} D{D}: " + "  ".join(
        f"pingpong {pp}: " + "/".join(f"{fl / ms / 1e9:.0f}" for ms in res[pp]) + " TF" for pp in (0, 1))
    print(line, flush=True)
    out.write(line + "\n")
fwb200.lib.fwb_attn_set_tuning(1002)
fwb200.lib.fwb_attn_set_tuning(200)
out.close()
