"""Target for `ncu --set full`: the dominant kernel (DiT self-attention, head_dim 128) at the BASELINE C2 size.
    ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/attn_d128 python tools/ncu_attn.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "fantasy-world_b200"))
import torch
import fwb200

kind = sys.argv[1] if len(sys.argv) > 1 else "attn128"
torch.manual_seed(0)
if kind == "attn128":
    B, H, L, D = 1, 40, 32760, 128
    q, k, v = (torch.randn(B, L, H, D, device="cuda").to(torch.bfloat16) for _ in range(3))
    for _ in range(3):
        fwb200.attention(q, k, v)
elif kind == "gemm":
    M, N, K = 32760, 5120, 5120
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    for _ in range(3):
        fwb200.linear(x, w, bias=b)
elif kind == "ln":                      # DiT pre-attention / pre-FFN LayerNorm + modulation: fp32 residual stream -> bf16
    rows, C = 32760, 5120
    x = torch.randn(rows, C, device="cuda")
    mul, add = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    for _ in range(3):
        fwb200.ln_modulate(x, eps=1e-6, mul=mul, add=add)
elif kind == "rms":                     # q / k RMSNorm + RoPE, in place on bf16
    rows, C, hd = 32760, 5120, 128
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(C, device="cuda")
    cs = torch.randn(rows, hd // 2, 2, device="cuda")
    for _ in range(3):
        fwb200.rmsnorm_rope_(x, w=w, eps=1e-6, cos_sin=cs, head_dim=hd)
torch.cuda.synchronize()
