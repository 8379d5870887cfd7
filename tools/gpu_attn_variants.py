"""Isolated timing of the attention kernel variants (fwb_attn_set_tuning 201 = v1 aliased S/P, 202 = decoupled attn2) at the hot
path's shapes, with torch SDPA (cuDNN / flash) beside them.  Run under gpurun; writes gpurun_out/attn_variants.log."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

import torch
import fwb200


def timeit(fn, iters=5, warm=2):
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


CASES = [(1, 40, 32760, 32760, 128), (1, 12, 32760, 32865, 96), (1, 16, 32865, 32865, 64), (21, 16, 1565, 1565, 64),
         (1, 40, 4095, 8190, 128), (1, 40, 32760, 512, 128)]
variants = [1, 2]
offsets = [int(a) for a in sys.argv[1:]] or [0, 1]
out = open(ROOT / "gpurun_out" / "attn_variants.log", "w")
for (B, H, Lq, Lk, D) in CASES:
    q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for L in (Lq, Lk, Lk))
    o = torch.empty_like(q)
    fl = 4.0 * B * H * Lq * Lk * D
    res, outs = [], {}
    sweep = []
    for off in offsets:
        fwb200.lib.fwb_attn_set_tuning(1000 + off)
        fwb200.lib.fwb_attn_set_tuning(202)
        sweep.append((off, timeit(lambda: fwb200.attention(q, k, v, out=o))))
    best_off = min(sweep, key=lambda r: r[1])[0]
    fwb200.lib.fwb_attn_set_tuning(1000 + best_off)
    for rep in range(2):
        for var in variants:
            fwb200.lib.fwb_attn_set_tuning(200 + var)
            ms = timeit(lambda: fwb200.attention(q, k, v, out=o))
            outs[var] = o.clone()
            res.append((var, ms))
    best = {var: min(ms for vv, ms in res if vv == var) for var in variants}
    qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
    ms_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt))
    ref = torch.nn.functional.scaled_dot_product_attention(qt, kt, vt).transpose(1, 2).float()
    errs = {var: (outs[var].float() - ref).abs().max().item() for var in variants}
    line = (f"B{B} H{H} Lq{Lq} Lk{Lk} D{D}: " + "  ".join(f"v{var} {best[var]:.3f} ms {fl / best[var] / 1e9:.0f} TF (|err| vs sdpa {errs[var]:.2e})"
                                                        for var in variants) + f"  | sdpa {ms_ref:.3f} ms {fl / ms_ref / 1e9:.0f} TF"
            + "  | v2 pingpong off/on: " + " ".join(f"{off}:{fl / ms / 1e9:.0f}" for off, ms in sweep))
    print(line, flush=True)
    out.write(line + "\n")
fwb200.lib.fwb_attn_set_tuning(200)
out.close()
