"""GPU bring-up harness (run under gpurun).  Every case runs in its own subprocess so that a faulting
kernel (sticky CUDA error) cannot take the remaining cases down.  Writes gpurun_out/bringup.log.

    python tools/gpu_bringup.py            # all cases
    python tools/gpu_bringup.py --case X   # one case, in-process
"""
from __future__ import annotations

import argparse
import math
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))


def _mma_case(N, K, a_tmem, b_mn, overrides=None):
    import torch
    import fwb200
    torch.manual_seed(0)
    A = torch.randn(128, K, device="cuda").to(torch.bfloat16)
    if b_mn:
        Bm = torch.randn(K, N, device="cuda").to(torch.bfloat16)
        ref = A.float() @ Bm.float()
    else:
        Bm = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        ref = A.float() @ Bm.float().t()
    D = fwb200.bringup_mma(A, Bm, N, K, a_tmem, b_mn, overrides)
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    print(f"max_abs_err={err:.4e} ref_absmax={ref.abs().max().item():.3f} ok={err < 1e-2}")


def _gemm_case(M, N, K, mode):
    import torch
    import fwb200
    torch.manual_seed(1)
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ref = x.float() @ w.float().t() + bias
    kw = dict(bias=bias)
    if mode == "plain":
        out = fwb200.linear(x, w, **kw)
    elif mode == "f32":
        out = fwb200.linear(x, w, out_dtype=torch.float32, **kw)
    elif mode == "gelu_tanh":
        out = fwb200.linear(x, w, act=fwb200.ACT_GELU_TANH, **kw)
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    elif mode == "gelu_erf":
        out = fwb200.linear(x, w, act=fwb200.ACT_GELU_ERF, out_dtype=torch.float32, **kw)
        ref = torch.nn.functional.gelu(ref)
    elif mode == "gate_resid":
        gate = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        out = fwb200.linear(x, w, scale1=gate, resid=resid, **kw)
        ref = resid.float() + gate * ref
    elif mode == "affine_f32resid":
        s1 = torch.randn(N, device="cuda"); t1 = torch.randn(N, device="cuda"); s2 = torch.randn(N, device="cuda")
        resid = torch.randn(M, N, device="cuda")
        out = fwb200.linear(x, w, scale1=s1, shift1=t1, scale2=s2, resid=resid, out_dtype=torch.float32, **kw)
        ref = resid + s2 * (s1 * ref + t1)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    tol = 5e-2 if out.dtype == torch.bfloat16 else 5e-3
    print(f"max_abs_err={err:.4e} ref_absmax={ref.abs().max().item():.3f} ok={err < tol * max(1.0, ref.abs().max().item())}")


def _attn_case(B, H, Lq, Lk, D, packed=False):
    import torch
    import fwb200
    torch.manual_seed(2)
    if packed:  # q,k,v are strided views of one [B, L, 3, H, D] buffer (VGGT qkv layout)
        assert Lq == Lk
        qkv = torch.randn(B, Lq, 3, H, D, device="cuda").to(torch.bfloat16)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    else:
        q = torch.randn(B, Lq, H, D, device="cuda").to(torch.bfloat16)
        k = torch.randn(B, Lk, H, D, device="cuda").to(torch.bfloat16)
        v = torch.randn(B, Lk, H, D, device="cuda").to(torch.bfloat16)
    out = fwb200.attention(q, k, v)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) / math.sqrt(D)
    ref = (torch.softmax(s, dim=-1) @ vf).permute(0, 2, 1, 3)
    err = (out.float() - ref).abs().max().item()
    print(f"max_abs_err={err:.4e} ref_absmax={ref.abs().max().item():.3f} nan={bool(torch.isnan(out.float()).any())} ok={err < 2e-2}")


def _time_case(kind):
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

    if kind == "gemm":
        for (M, N, K) in [(32760, 5120, 5120), (32760, 13824, 5120), (32760, 5120, 13824), (32865, 3072, 1024),
                          (32865, 4096, 1024), (32865, 1024, 4096)]:
            x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
            bias = torch.randn(N, device="cuda")
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            res = []
            for mode in (1, 2):
                fwb200.lib.fwb_gemm_set_mode(mode)
                ms = timeit(lambda: fwb200.linear(x, w, bias=bias, out=out))
                res.append(f"mode{mode} {ms:.3f} ms {2*M*N*K/ms/1e9:.0f} TF")
            fwb200.lib.fwb_gemm_set_mode(2)
            ms = timeit(lambda: fwb200.linear(x, w, bias=bias, out=out, act=fwb200.ACT_GELU_ERF, round_flags=3))
            res.append(f"mode2+gelu_erf {ms:.3f} ms")
            fwb200.lib.fwb_gemm_set_mode(-1)
            ms_ref = timeit(lambda: torch.nn.functional.linear(x, w))
            print(f"gemm M={M} N={N} K={K}: " + " | ".join(res) + f" | cublas {ms_ref:.3f} ms {2*M*N*K/ms_ref/1e9:.0f} TF")
    else:
        for (B, H, L, D) in [(1, 40, 32760, 128), (1, 16, 32865, 64), (1, 12, 32760, 96), (21, 16, 1565, 64)]:
            q = torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16)
            k = torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16)
            v = torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16)
            out = torch.empty_like(q)
            fl = 4 * B * H * L * L * D
            res = []
            for poly in (0, 2):
                fwb200.lib.fwb_attn_set_exp2_poly(poly)
                ms = timeit(lambda: fwb200.attention(q, k, v, out=out), iters=3, warm=1)
                res.append(f"poly{poly}/8 {ms:.3f} ms {fl/ms/1e9:.0f} TF")
            fwb200.lib.fwb_attn_set_exp2_poly(-1)
            print("   variants: " + " | ".join(res))
            ms = timeit(lambda: fwb200.attention(q, k, v, out=out), iters=3, warm=1)
            qt, kt, vt = (t.transpose(1, 2) for t in (q, k, v))
            ms_ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), iters=3, warm=1)
            print(f"attn B={B} H={H} L={L} D={D}: ours {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s | sdpa {ms_ref:.3f} ms {fl/ms_ref/1e9:.1f} TFLOP/s")


CASES = {
    # name: (fn, args)
    "mma_ss_kk_n128_k64": (_mma_case, (128, 64, False, False)),
    "mma_ss_kk_n256_k128": (_mma_case, (256, 128, False, False)),
    "mma_ss_kk_n64_k256": (_mma_case, (64, 256, False, False)),
    "mma_ss_mn_n128_k64": (_mma_case, (128, 64, False, True)),
    "mma_ss_mn_n128_k128": (_mma_case, (128, 128, False, True)),
    "mma_ss_mn_n64_k128": (_mma_case, (64, 128, False, True)),
    "mma_ss_mn_n128_k128_swapped": (_mma_case, (128, 128, False, True, [None, None, None, 1024, 128 * 128, None, None, None])),
    "mma_ts_kk_n128_k128": (_mma_case, (128, 128, True, False)),
    "mma_ts_mn_n128_k128": (_mma_case, (128, 128, True, True)),
    "mma_ts_mn_n64_k128": (_mma_case, (64, 128, True, True)),
    "mma_ts_kk_n128_k128_adv4": (_mma_case, (128, 128, True, False, [None] * 7 + [4])),
    "mma_ts_kk_n128_k128_adv16": (_mma_case, (128, 128, True, False, [None] * 7 + [16])),
    "gemm_small_plain": (_gemm_case, (256, 256, 128, "plain")),
    "gemm_ragged_plain": (_gemm_case, (1000, 1152, 1024, "plain")),
    "gemm_f32": (_gemm_case, (1565, 1024, 4096, "f32")),
    "gemm_gelu_tanh": (_gemm_case, (777, 13824, 5120, "gelu_tanh")),
    "gemm_gelu_erf": (_gemm_case, (777, 4096, 1024, "gelu_erf")),
    "gemm_gate_resid": (_gemm_case, (4095, 5120, 5120, "gate_resid")),
    "gemm_affine_f32resid": (_gemm_case, (1565, 1024, 4096, "affine_f32resid")),
    "gemm_n64": (_gemm_case, (1000, 64, 5120, "plain")),
    "attn_d128_small": (_attn_case, (1, 2, 256, 256, 128)),
    "attn_d128_ragged": (_attn_case, (1, 3, 1000, 777, 128)),
    "attn_d128_long": (_attn_case, (1, 2, 4095, 8190, 128)),
    "attn_d64_small": (_attn_case, (2, 4, 300, 300, 64)),
    "attn_d64_packed": (_attn_case, (3, 16, 1565, 1565, 64, True)),
    "attn_d96": (_attn_case, (1, 12, 1560, 1565, 96)),
    "attn_d128_tiny": (_attn_case, (1, 1, 16, 21, 128)),
    "time_gemm": (_time_case, ("gemm",)),
    "time_attn": (_time_case, ("attn",)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if args.case:
        fn, a = CASES[args.case]
        fn(*a)
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    log = open(out_dir / "bringup.log", "w")
    for name in CASES:
        if args.only and not any(name.startswith(p) for p in args.only.split(",")):
            continue
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=300)
            tail = (r.stdout.strip().splitlines() or ["<no stdout>"])
            err_tail = r.stderr.strip().splitlines()[-3:] if r.returncode != 0 else []
            line = f"[{name}] rc={r.returncode} {time.time()-t0:.1f}s :: " + " | ".join(tail[-8:]) + (" :: ERR " + " / ".join(err_tail) if err_tail else "")
        except subprocess.TimeoutExpired:
            line = f"[{name}] TIMEOUT"
        print(line, flush=True)
        log.write(line + "\n")
        log.flush()


if __name__ == "__main__":
    main()
