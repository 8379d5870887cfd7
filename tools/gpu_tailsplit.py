"""Isolated timing of the attention tile schedule with / without the tail split at the per-rank shapes of the 8-GPU
sequence-parallel run (and the 1-GPU shapes as a control).  Run under gpurun; writes gpurun_out/tailsplit.log."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

import torch
import fwb200


def timeit(fn, iters=10, warm=3):
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


CASES = [
    ("dit self, 8 ranks, 1 of 4 K|V slices", 1, 40, 4095, 8190, 128, False),
    ("dit self, 8 ranks, 1 of 4 K|V slices, partial", 1, 40, 4095, 8190, 128, True),
    ("dit self, 8 ranks, all keys", 1, 40, 4095, 32760, 128, False),
    ("dit self, 4 ranks, 1 of 4 slices", 1, 40, 8190, 8190, 128, True),
    ("adapter video<-geo, 8 ranks", 1, 12, 4095, 32865, 96, False),
    ("adapter geo<-video, 8 ranks", 1, 12, 4695, 32760, 96, False),
    ("vggt global, 8 ranks", 1, 16, 4695, 32865, 64, False),
    ("vggt frame, 3-frame shard", 3, 16, 1565, 1565, 64, False),
    ("adapter video<-geo, 1 GPU (control)", 1, 12, 32760, 32865, 96, False),
    ("dit self, 1 GPU (control)", 1, 40, 32760, 32760, 128, False),
]

out = open(ROOT / "gpurun_out" / "tailsplit.log", "w")
for name, B, H, Lq, Lk, D, partial in CASES:
    q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for L in (Lq, Lk, Lk))
    o = torch.empty_like(q)
    part = torch.empty(B, Lq, H, D, device="cuda")
    lse = torch.empty(B, H, Lq, device="cuda")
    fn = (lambda: fwb200.attention_partial(q, k, v, part, lse)) if partial else (lambda: fwb200.attention(q, k, v, out=o))
    res = []
    for mode in (100, 101, 100, 101):
        fwb200.lib.fwb_attn_set_tuning(mode)
        res.append(timeit(fn))
    fwb200.lib.fwb_attn_set_tuning(101)
    fl = 4.0 * B * H * Lq * Lk * D
    off, on = min(res[0], res[2]), min(res[1], res[3])
    line = (f"{name}: B{B} H{H} Lq{Lq} Lk{Lk} D{D} tiles={((Lq + 255) // 256) * H * B}  unsplit {off:.3f} ms ({fl / off / 1e9:.0f} TF)  "
            f"tail-split {on:.3f} ms ({fl / on / 1e9:.0f} TF)  speedup {off / on:.3f}")
    print(line, flush=True)
    out.write(line + "\n")
out.close()
