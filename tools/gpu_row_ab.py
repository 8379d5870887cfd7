"""Timing of the persistent row kernels on the GPU at the in-step shapes, by CTAs per SM (fwb_rowwise_set_ctas_per_sm; 0 = automatic).
CUDA-event times over many launches at the in-step shapes; inputs (671 MB .. 1 GB per launch) exceed the 126 MB L2.

    python tools/gpu_row_ab.py            # prints TB/s per configuration (algorithmic bytes: read + write once)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "fantasy-world_b200"))
import torch
import fwb200


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    torch.manual_seed(0)
    rows, C, hd = 32760, 5120, 128
    xf = torch.randn(rows, C, device="cuda")
    xb = xf.to(torch.bfloat16)
    mul, add, w = torch.randn(C, device="cuda"), torch.randn(C, device="cuda"), torch.rand(C, device="cuda") + 0.5
    cs = torch.randn(rows, hd // 2, 2, device="cuda")
    rows_v, Cv = 81 * 1565, 1024                      # VGGT token rows
    xv = torch.randn(rows_v, Cv, device="cuda")
    wv, bv = torch.randn(Cv, device="cuda"), torch.randn(Cv, device="cuda")
    cases = [
        ("ln_modulate fp32->bf16 C5120", lambda: fwb200.ln_modulate(xf, eps=1e-6, mul=mul, add=add), rows * C * 6),
        ("ln_modulate bf16->bf16 C5120", lambda: fwb200.ln_modulate(xb, eps=1e-6, mul=mul, add=add), rows * C * 4),
        ("rmsnorm_rope bf16 in place C5120", lambda: fwb200.rmsnorm_rope_(xb, w=w, eps=1e-6, cos_sin=cs, head_dim=hd), rows * C * 4),
        ("ln_modulate fp32->bf16 C1024 (VGGT)", lambda: fwb200.ln_modulate(xv, eps=1e-5, w=wv, b=bv, mul=wv, add=bv), rows_v * Cv * 6),
    ]
    for name, fn, nbytes in cases:
        line = [f"{name:40s}"]
        for ctas in (0, 2, 3, 4, 6, 8):
            fwb200.lib.fwb_rowwise_set_ctas_per_sm(ctas)
            t = timeit(fn)
            line.append(f"cta{ctas}: {nbytes / t / 1e12:5.2f}")
        print(" | ".join(line), flush=True)
    fwb200.lib.fwb_rowwise_set_ctas_per_sm(0)


if __name__ == "__main__":
    main()
