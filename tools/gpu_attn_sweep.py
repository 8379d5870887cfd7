"""Interleaved timing of the attention kernel variants at the hot-path shapes (run under gpurun):

    python tools/gpu_attn_sweep.py [--quick]

For every shape: kernel variant (1 = S/P aliased, 2 = decoupled) x exp2-polynomial share (0 or 2 pairs of 8) x MUFU ping-pong, 3 interleaved rounds each (the box runs power-capped: +-4 % between rounds), plus cuDNN / flash SDPA of
torch on the same tensors.  Writes gpurun_out/r02_attn_sweep.log (TFLOP/s = 4 B H Lq Lk D / time)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
import torch
import fwb200

QUICK = "--quick" in sys.argv


def timeit(fn, iters=4, warm=1):
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


SHAPES = [(1, 40, 32760, 32760, 128, "DiT self"), (1, 12, 32760, 32865, 96, "adapter v<-g"), (1, 16, 32865, 32865, 64, "VGGT global"),
          (21, 16, 1565, 1565, 64, "VGGT frame"), (1, 40, 32760, 512, 128, "cross text"), (1, 40, 32760, 257, 128, "cross CLIP"), (1, 40, 8190, 32760, 128, "DiT self, 4-rank shard")]
if QUICK:
    SHAPES = SHAPES[:3]
log = open(ROOT / "gpurun_out" / "r02_attn_sweep.log", "w")


def emit(line):
    print(line, flush=True)
    log.write(line + "\n")
    log.flush()


lib = fwb200.lib
for (B, H, Lq, Lk, D, name) in SHAPES:
    q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for L in (Lq, Lk, Lk))
    o = torch.empty_like(q)
    fl = 4.0 * B * H * Lq * Lk * D
    configs = [(1, 0, 0), (1, 2, 0), (1, 2, 1), (2, 0, 1), (2, 2, 1), (2, 2, 0)]
    res = {c: [] for c in configs}
    for rnd in range(3):
        for c in configs:
            var, poly, pp = c
            lib.fwb_attn_set_variant(var)
            lib.fwb_attn_set_exp2_poly(poly)
            if var in (1, 2):
                lib.fwb_attn_set_mufu_pingpong(var, pp)
            res[c].append(fl / timeit(lambda: fwb200.attention(q, k, v, out=o)) / 1e9)
    if D in (96, 128) and (Lq + 255) // 256 % 2 == 0:     # CTA pairs sharing K/V tiles by TMA multicast, A/B on the default kernel
        lib.fwb_attn_set_variant(0)
        lib.fwb_attn_set_exp2_poly(-1)
        ab = {0: [], 1: []}
        for rnd in range(4):
            for on in (0, 1):
                lib.fwb_attn_set_multicast(on)
                ab[on].append(fl / timeit(lambda: fwb200.attention(q, k, v, out=o), iters=6) / 1e9)
        lib.fwb_attn_set_multicast(1)
        emit(f"   [{name}] multicast off: " + "/".join(f"{x:.0f}" for x in ab[0]) + "   on: " + "/".join(f"{x:.0f}" for x in ab[1]) + " TF")
    if Lk <= 2048:     # short-key configuration (one Q tile per CTA, two CTAs per SM) A/B under the default policy
        lib.fwb_attn_set_variant(0)
        lib.fwb_attn_set_exp2_poly(-1)
        ab = {0: [], 2048: []}
        for rnd in range(4):
            for mx in (0, 2048):
                lib.fwb_attn_set_short_kv_max(mx)
                ab[mx].append(fl / timeit(lambda: fwb200.attention(q, k, v, out=o), iters=6) / 1e9)
        lib.fwb_attn_set_short_kv_max(2048)
        emit(f"   [{name}] short-kv config off: " + "/".join(f"{x:.0f}" for x in ab[0]) + "   on: " + "/".join(f"{x:.0f}" for x in ab[2048]) + " TF")
    if D == 96:     # native PV width A/B on the default kernel
        lib.fwb_attn_set_variant(0)
        lib.fwb_attn_set_exp2_poly(-1)
        ab = {0: [], 1: []}
        for rnd in range(3):
            for on in (0, 1):
                lib.fwb_attn_set_pv_n96(on)
                ab[on].append(fl / timeit(lambda: fwb200.attention(q, k, v, out=o)) / 1e9)
        lib.fwb_attn_set_pv_n96(1)
        emit("   head_dim 96 PV N=128: " + "/".join(f"{x:.0f}" for x in ab[0]) + "   PV N=96: " + "/".join(f"{x:.0f}" for x in ab[1]) + " TF")
    lib.fwb_attn_set_variant(0)
    lib.fwb_attn_set_exp2_poly(-1)
    lib.fwb_attn_set_mufu_pingpong(1, 0)
    lib.fwb_attn_set_mufu_pingpong(2, 1)
    qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
    sd = [fl / timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt)) / 1e9 for _ in range(3)]
    emit(f"== {name}: B{B} H{H} Lq{Lq} Lk{Lk} D{D}   torch SDPA: " + "/".join(f"{x:.0f}" for x in sd) + " TF")
    for c in sorted(configs, key=lambda c: -sorted(res[c])[1]):
        emit(f"   variant {c[0]} poly {c[1]}/8 pingpong {c[2]}: " + "/".join(f"{x:.0f}" for x in res[c]) + f" TF  (median {sorted(res[c])[1]:.0f})")
log.close()
