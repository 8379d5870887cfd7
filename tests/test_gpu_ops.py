"""GPU parity tests of the individual C-ABI kernels (called through ctypes — fwb200.ops) against plain fp32 torch math on
the same bf16-rounded inputs.  Tolerances are written per test; bf16 outputs carry one rounding (2^-9 relative)."""
import math

import pytest
import torch

from _common import assert_close_frac

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fwb():
    import fwb200
    fwb200.require_device()
    return fwb200


def _bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------------------------
# GEMM + epilogues
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 5120, 256), (128, 256, 64), (1000, 1152, 1024), (777, 13824, 5120), (1565, 1024, 4096),
                                   (300, 64, 5120), (4095, 5120, 144), (257, 2304, 5120)])
def test_gemm_fp32_out_matches_fp32_reference(fwb, M, N, K):
    torch.manual_seed(M + N + K)
    x = _bf(torch.randn(M, K, device="cuda"))
    w = _bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    b = torch.randn(N, device="cuda")
    out = fwb.linear(x, w, bias=b, out_dtype=torch.float32)
    ref = x.float() @ w.float().t() + b
    # fp32 accumulation in a different order: rtol 1e-3 / atol 1e-4 (the north-star tolerance) holds for fp32 outputs
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-4 * math.sqrt(K / 64))


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1000, 512, 1024), (4095, 5120, 5120), (32760 // 8 + 77, 1024, 4096)])
def test_gemm_kernel_variants_agree(fwb, mode, M, N, K):
    """single-CTA 128x128 / 128x256 and the CTA-pair (cta_group::2) 256x256 kernels compute the same GEMM."""
    torch.manual_seed(M + N + K)
    x = _bf(torch.randn(M, K, device="cuda"))
    w = _bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda")
    try:
        fwb.lib.fwb_gemm_set_mode(mode)
        out = fwb.linear(x, w, bias=b, act=fwb.ACT_GELU_ERF, resid=r, out_dtype=torch.float32)
    finally:
        fwb.lib.fwb_gemm_set_mode(-1)
    ref = torch.nn.functional.gelu(x.float() @ w.float().t() + b) + r
    torch.testing.assert_close(out, ref, rtol=1e-3, atol=1e-4 * math.sqrt(K / 64))


def test_gemm_epilogue_family(fwb):
    torch.manual_seed(3)
    M, N, K = 515, 1024, 1024
    x = _bf(torch.randn(M, K, device="cuda"))
    w = _bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    b, s1, t1, s2 = (torch.randn(N, device="cuda") for _ in range(4))
    lin = x.float() @ w.float().t() + b
    r16 = _bf(torch.randn(M, N, device="cuda"))
    r32 = torch.randn(M, N, device="cuda")

    def rb(t):
        return t.to(torch.bfloat16).float()

    out = fwb.linear(x, w, bias=b, act=fwb.ACT_GELU_TANH, round_flags=fwb.ROUND_AFTER_BIAS | fwb.ROUND_AFTER_ACT)
    ref = rb(torch.nn.functional.gelu(rb(lin), approximate="tanh"))
    assert_close_frac(out.float(), ref, rtol=2e-2, atol=2e-2, loose_atol=1e-1)   # tanh.approx + 1 bf16 ulp
    out = fwb.linear(x, w, bias=b, act=fwb.ACT_GELU_ERF, out_dtype=torch.float32)
    torch.testing.assert_close(out, torch.nn.functional.gelu(lin), rtol=1e-3, atol=1e-4)
    out = fwb.linear(x, w, bias=b, act=fwb.ACT_RELU, out_dtype=torch.float32)
    torch.testing.assert_close(out, torch.relu(lin), rtol=1e-3, atol=1e-4)
    out = fwb.linear(x, w, bias=b, act=fwb.ACT_SILU, out_dtype=torch.float32)
    torch.testing.assert_close(out, torch.nn.functional.silu(lin), rtol=1e-3, atol=1e-4)
    # DiT gate + residual: bf16(x + bf16(gate * bf16(lin)))
    out = fwb.linear(x, w, bias=b, scale1=s1, resid=r16, round_flags=fwb.ROUND_AFTER_BIAS | fwb.ROUND_AFTER_AFFINE)
    ref = rb(r16.float() + rb(s1 * rb(lin)))
    assert_close_frac(out.float(), ref, rtol=1.6e-2, atol=1e-2, loose_atol=2e-1)
    # VGGT fc2: x32 + s2 * (s1 * bf16(lin) + t1)
    out = fwb.linear(x, w, bias=b, scale1=s1, shift1=t1, scale2=s2, resid=r32, out_dtype=torch.float32, round_flags=fwb.ROUND_AFTER_BIAS)
    ref = r32 + s2 * (s1 * rb(lin) + t1)
    assert_close_frac(out, ref, rtol=1e-3, atol=2e-3, loose_atol=2e-1)     # a bf16 tie in rb(lin) moves the result by s1*s2*ulp
    # strided A (a column slice of a wider buffer) and in-place residual
    big = _bf(torch.randn(M, 2 * K, device="cuda"))
    out = fwb.linear(big[:, K:], w, out_dtype=torch.float32)
    torch.testing.assert_close(out, big[:, K:].float() @ w.float().t(), rtol=1e-3, atol=1e-3)


def test_gemm_argument_errors(fwb):
    x = _bf(torch.randn(16, 20, device="cuda"))
    w = _bf(torch.randn(8, 20, device="cuda"))
    with pytest.raises(RuntimeError, match="multiples of 8"):
        fwb.linear(x, w)


# ------------------------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, scale=None):
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * (scale if scale is not None else 1 / math.sqrt(q.shape[-1]))
    return (torch.softmax(s, dim=-1) @ vf).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,H,Lq,Lk,D", [(1, 2, 256, 256, 128), (1, 3, 1000, 777, 128), (2, 4, 300, 300, 64), (1, 12, 1560, 1565, 96),
                                         (1, 1, 16, 21, 128), (1, 2, 1, 1, 64), (1, 2, 129, 1, 128), (3, 16, 1565, 1565, 64),
                                         (1, 40, 520, 257, 128), (1, 2, 4095, 8190, 128)])
def test_attention_matches_fp32_softmax(fwb, B, H, Lq, Lk, D):
    torch.manual_seed(B * 1000 + Lq + Lk + D)
    q, k, v = (_bf(torch.randn(B, L, H, D, device="cuda")) for L in (Lq, Lk, Lk))
    out = fwb.attention(q, k, v)
    ref = _attn_ref(q, k, v)
    # P is rounded to bf16 before PV (as in flash-attn / cuDNN) and the output is bf16: 2^-8 relative on O(1) values
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=6e-3)
    assert not torch.isnan(out.float()).any()


def test_attention_large_logits_and_lazy_rescale(fwb):
    """Row max that keeps growing tile after tile (sorted keys) exercises the O-rescale path; large |scores| exercise exp2 range."""
    torch.manual_seed(0)
    B, H, L, D = 1, 2, 1024, 128
    q = _bf(torch.randn(B, L, H, D, device="cuda") * 3)
    k = _bf(torch.randn(B, L, H, D, device="cuda") * torch.linspace(0.1, 4, L, device="cuda").view(1, L, 1, 1))
    v = _bf(torch.randn(B, L, H, D, device="cuda"))
    out = fwb.attention(q, k, v)
    torch.testing.assert_close(out.float(), _attn_ref(q, k, v), rtol=3e-2, atol=2e-2)


def test_attention_strided_views_and_accumulate(fwb):
    torch.manual_seed(1)
    B, L, H, D = 2, 333, 16, 64
    qkv = _bf(torch.randn(B, L, 3, H, D, device="cuda"))
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    out = fwb.attention(q, k, v)
    torch.testing.assert_close(out.float(), _attn_ref(q, k, v), rtol=2e-2, atol=6e-3)
    # accumulate: out = bf16(out + bf16(second attention)) — the text + CLIP sum of the DiT cross attention
    k2, v2 = _bf(torch.randn(B, 257, H, D, device="cuda")), _bf(torch.randn(B, 257, H, D, device="cuda"))
    first = out.clone()
    fwb.attention(q, k2, v2, out=out, accumulate=True)
    ref = (first.float() + _attn_ref(q, k2, v2).to(torch.bfloat16).float())
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=1.2e-2)


def test_attention_split_kv_partials_and_merge(fwb):
    """Split-KV mode used by the sequence-parallel pipeline: partial attentions over disjoint key subsets + merge == one
    attention over all keys (up to fp32 re-association before the single bf16 rounding)."""
    torch.manual_seed(11)
    B, H, Lq, D = 1, 5, 700, 128
    sizes = [256, 900, 333]
    q = _bf(torch.randn(B, Lq, H, D, device="cuda"))
    k = _bf(torch.randn(B, sum(sizes), H, D, device="cuda") * 1.5)
    v = _bf(torch.randn(B, sum(sizes), H, D, device="cuda"))
    part = torch.empty(len(sizes), B, Lq, H, D, device="cuda")
    lse = torch.empty(len(sizes), B, H, Lq, device="cuda")
    o = 0
    for i, n in enumerate(sizes):
        fwb.attention_partial(q, k[:, o:o + n], v[:, o:o + n], part[i], lse[i])
        o += n
    merged = fwb.attention_merge(part, lse)
    full = fwb.attention(q, k, v)
    torch.testing.assert_close(merged.float(), _attn_ref(q, k, v), rtol=2e-2, atol=6e-3)
    assert (merged.float() - full.float()).abs().max() <= 2 ** -7 * full.float().abs().max()   # at most a bf16 ulp apart
    # lse is the base-2 log-sum-exp of the scaled scores
    s = (q.float().permute(0, 2, 1, 3) @ k[:, :sizes[0]].float().permute(0, 2, 3, 1)) / math.sqrt(D)
    ref_lse = torch.logsumexp(s, dim=-1) / math.log(2)
    torch.testing.assert_close(lse[0], ref_lse, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("B,H,Lq,Lk,D", [(1, 40, 4095, 8190, 128),      # DiT self-attention shard at 8 ranks, one K|V slice: 640 tiles
                                         (1, 12, 4095, 32865, 96),      # adapter, video <- geometry, 8 ranks: 192 tiles
                                         (1, 16, 4695, 32865, 64),      # VGGT global attention shard: 304 tiles
                                         (3, 16, 1565, 1565, 64),       # VGGT frame attention of a 3-frame shard: 336 tiles
                                         (1, 2, 300, 5000, 128)])       # fewer tiles than SMs
def test_attention_tail_split_matches_unsplit(fwb, B, H, Lq, Lk, D):
    """Tile schedule: the last, partly filled wave of (256-row x head) tiles is split along the keys over the idle SMs and
    merged.  The result must agree with the unsplit schedule to one bf16 rounding, in the bf16-output and the split-KV
    (fp32 partial + lse) modes alike."""
    torch.manual_seed(Lq + Lk + D)
    q, k, v = (_bf(torch.randn(B, L, H, D, device="cuda")) for L in (Lq, Lk, Lk))
    part = torch.empty(2, B, Lq, H, D, device="cuda")
    lse = torch.empty(2, B, H, Lq, device="cuda")
    try:
        fwb.lib.fwb_attn_set_tail_split(0)
        ref = fwb.attention(q, k, v)
        fwb.attention_partial(q, k, v, part[0], lse[0])
    finally:
        fwb.lib.fwb_attn_set_tail_split(1)
    out = fwb.attention(q, k, v)
    fwb.attention_partial(q, k, v, part[1], lse[1])
    torch.cuda.synchronize()
    scale = ref.float().abs().max()
    assert (out.float() - ref.float()).abs().max() <= 2 ** -7 * scale
    # P is rounded to bf16 relative to a (stale-tolerant) running row max that differs between the schedules: the fp32 partials
    # carry that rounding noise (2^-9 per element of P, a few sigma of the resulting sum), far below a bf16 ulp of the output
    assert (part[1] - part[0]).abs().max() <= 2 ** -8 * scale
    torch.testing.assert_close(lse[1], lse[0], rtol=0, atol=2e-3)
    # and against fp32 math on a slice of the rows (the full score matrix would not fit for the large cases)
    rows = slice(Lq - 300, Lq)
    torch.testing.assert_close(out[:, rows].float(), _attn_ref(q[:, rows], k, v), rtol=2e-2, atol=6e-3)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("B,H,Lq,Lk,D", [(1, 3, 1000, 777, 128), (2, 4, 300, 1300, 64), (1, 12, 1560, 1565, 96), (1, 2, 129, 1, 128),
                                         (1, 2, 257, 65, 128), (1, 2, 300, 128, 64),
                                         (1, 8, 4095, 8190, 128), (3, 16, 1565, 1565, 64)])
def test_attention_kernel_variants(fwb, variant, B, H, Lq, Lk, D):
    """Both attention kernels (1: P aliased on S, one MMA thread; 2: decoupled S / P, one MMA warp per Q tile, MUFU ping-pong)
    against fp32 math, including the split-KV outputs.  The default picks one per head_dim; both stay covered."""
    torch.manual_seed(variant * 7 + Lq + Lk + D)
    q, k, v = (_bf(torch.randn(B, L, H, D, device="cuda")) for L in (Lq, Lk, Lk))
    part = torch.empty(1, B, Lq, H, D, device="cuda")
    lse = torch.empty(1, B, H, Lq, device="cuda")
    try:
        fwb.lib.fwb_attn_set_variant(variant)
        out = fwb.attention(q, k, v)
        fwb.attention_partial(q, k, v, part[0], lse[0])
        merged = fwb.attention_merge(part, lse)
        torch.cuda.synchronize()
    finally:
        fwb.lib.fwb_attn_set_variant(0)
    ref = _attn_ref(q, k, v)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=6e-3)
    torch.testing.assert_close(merged.float(), ref, rtol=2e-2, atol=6e-3)
    s = (q.float().permute(0, 2, 1, 3) @ k.float().permute(0, 2, 3, 1)) / math.sqrt(D)
    torch.testing.assert_close(lse[0], torch.logsumexp(s, dim=-1) / math.log(2), rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("poly", [2])
@pytest.mark.parametrize("B,H,Lq,Lk,D", [(1, 3, 1000, 3000, 128), (1, 12, 1560, 1565, 96), (2, 4, 300, 1300, 64)])
def test_attention_exp2_polynomial_share(fwb, variant, poly, B, H, Lq, Lk, D):
    """fwb_attn_set_exp2_poly: `poly` of every 8 softmax element pairs take 2^x from the packed FMA-pipe polynomial (max rel. error
    8.6e-5) instead of MUFU.EX2.  P is rounded to bf16 right after (rel. 2e-3), so the output may move by a fraction of a bf16 ulp
    only: compared with the MUFU-only kernel at 2^-7 of the output scale, and with fp32 math at the usual attention tolerance.
    Keys are scaled up along the sequence so that the rescale path (running max growing by > 2^8) is exercised too."""
    torch.manual_seed(5)
    q = _bf(torch.randn(B, Lq, H, D, device="cuda") * 2)
    k = _bf(torch.randn(B, Lk, H, D, device="cuda") * torch.linspace(0.2, 3, Lk, device="cuda").view(1, Lk, 1, 1))
    v = _bf(torch.randn(B, Lk, H, D, device="cuda"))
    try:
        fwb.lib.fwb_attn_set_variant(variant)
        fwb.lib.fwb_attn_set_exp2_poly(0)
        ref = fwb.attention(q, k, v)
        fwb.lib.fwb_attn_set_exp2_poly(poly)
        out = fwb.attention(q, k, v)
        torch.cuda.synchronize()
    finally:
        fwb.lib.fwb_attn_set_exp2_poly(-1)
        fwb.lib.fwb_attn_set_variant(0)
    assert (out.float() - ref.float()).abs().max() <= 2 ** -7 * ref.float().abs().max()
    torch.testing.assert_close(out.float(), _attn_ref(q, k, v), rtol=2e-2, atol=6e-3)


def test_attention_softmax_rows_sum_to_one_full_size(fwb):
    """Size-independent property at the BASELINE C2 size (L = 32760 tokens, 128-dim heads): with V = 1 the output is 1."""
    torch.manual_seed(2)
    B, H, L, D = 1, 2, 32760, 128
    q, k = _bf(torch.randn(B, L, H, D, device="cuda")), _bf(torch.randn(B, L, H, D, device="cuda"))
    v = torch.ones(B, L, H, D, device="cuda", dtype=torch.bfloat16)
    out = fwb.attention(q, k, v)
    assert (out.float() - 1).abs().max() < 8e-3
    # linearity in V: attn(q,k,2v) == 2 attn(q,k,v) exactly (power-of-two scaling commutes with every rounding)
    v = _bf(torch.randn(B, L, H, D, device="cuda"))
    assert torch.equal(fwb.attention(q, k, (2 * v.float()).to(torch.bfloat16)).float(), 2 * fwb.attention(q, k, v).float())


# ------------------------------------------------------------------------------------------------------------------
# row kernels
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,C,dtype", [(37, 5120, torch.bfloat16), (50, 1024, torch.float32), (9, 1280, torch.bfloat16),
                                          (5, 2048, torch.float32), (3, 2560, torch.bfloat16)])
def test_ln_modulate(fwb, rows, C, dtype):
    torch.manual_seed(rows + C)
    x = (torch.randn(rows, C, device="cuda") * 2 + 0.3).to(dtype)
    w, b, mul, add = (torch.randn(C, device="cuda") for _ in range(4))
    for kw in (dict(), dict(w=w, b=b), dict(mul=mul, add=add), dict(w=w, b=b, mul=mul, add=add)):
        out = fwb.ln_modulate(x, eps=1e-6, **kw)
        y = torch.nn.functional.layer_norm(x.float(), (C,), kw.get("w"), kw.get("b"), 1e-6)
        if "mul" in kw:
            y = y * mul + add
        torch.testing.assert_close(out.float(), y.to(torch.bfloat16).float(), rtol=8e-3, atol=1e-2)
        assert out.dtype == torch.bfloat16


def test_rmsnorm_rope(fwb):
    from oracle import fw_oracle as O
    torch.manual_seed(5)
    f, h, w, H, D = 2, 3, 4, 40, 128
    L, C = f * h * w, H * D
    x = _bf(torch.randn(L, C, device="cuda"))
    wt = torch.randn(C, device="cuda").abs() + 0.5
    tab = O.rope_table_3d(D, f, h, w)
    cs = torch.stack([tab.real, tab.imag], -1).float().cuda().contiguous()
    y = x.clone()
    fwb.rmsnorm_rope_(y, w=wt, eps=1e-6, cos_sin=cs, head_dim=D)
    xn = O.rms_norm(x.float().cpu()[None], wt.cpu(), 1e-6, O.BF16)
    ref = O.rope_apply(xn, tab, H, O.BF16)[0]
    assert_close_frac(y.float().cpu(), ref, rtol=1.6e-2, atol=1.6e-2, loose_atol=1.3e-1)   # 1 bf16 ulp of the O(1..8) inputs feeding the rotation
    # rope only, head_dim 96, strided rows (the adapter's [q | v] buffer)
    buf = _bf(torch.randn(L, 2304, device="cuda"))
    tab96 = O.rope_table_3d(96, f, h, w)
    cs96 = torch.stack([tab96.real, tab96.imag], -1).float().cuda().contiguous()
    orig = buf.clone()
    fwb.rmsnorm_rope_(buf[:, :1152], cos_sin=cs96, head_dim=96)
    ref = O.rope_apply(orig[:, :1152].float().cpu()[None], tab96, 12, O.BF16)[0]
    assert_close_frac(buf[:, :1152].float().cpu(), ref, rtol=1.6e-2, atol=1.6e-2, loose_atol=1.3e-1)
    assert torch.equal(buf[:, 1152:], orig[:, 1152:])


def test_ln64_rope2d(fwb):
    import fwb200.engine as E
    from oracle import fw_oracle as O
    torch.manual_seed(6)
    S, hh, ww, H = 2, 3, 5, 16
    P = 5 + hh * ww
    sd = {"a.camera_token": torch.zeros(1, 2, 1, 1024), "a.register_token": torch.zeros(1, 2, 4, 1024)}
    _, pos = O.aggregator_input(sd, "a", torch.zeros(1, S, hh, ww, 1024))
    qkv = _bf(torch.randn(S * P, 3 * H * 64, device="cuda"))
    qw, qb, kw, kb = (torch.randn(64, device="cuda") for _ in range(4))
    E.ROPE2D_FP32_ANGLES = True   # compare with the fp32-angle oracle
    try:
        cosT, sinT = E.rope2d_expanded(pos.cuda())
    finally:
        E.ROPE2D_FP32_ANGLES = False
    y = qkv.clone()
    fwb.ln64_rope2d_(y, H, eps=1e-5, qw=qw, qb=qb, kw=kw, kb=kb, cosT=cosT, sinT=sinT)
    q5 = qkv.float().cpu().view(S, P, 3, H, 64).permute(2, 0, 3, 1, 4)
    q = O.rope2d_apply(O.layer_norm(q5[0], 1e-5, qw.cpu(), qb.cpu()), pos)
    k = O.rope2d_apply(O.layer_norm(q5[1], 1e-5, kw.cpu(), kb.cpu()), pos)
    got = y.float().cpu().view(S, P, 3, H, 64).permute(2, 0, 3, 1, 4)
    assert_close_frac(got[0], q.to(torch.bfloat16).float(), rtol=1.6e-2, atol=3e-2, loose_atol=1e-1)
    assert_close_frac(got[1], k.to(torch.bfloat16).float(), rtol=1.6e-2, atol=3e-2, loose_atol=1e-1)
    assert torch.equal(got[2], q5[2])  # V untouched


def test_cfg_euler_step(fwb):
    torch.manual_seed(7)
    n = 16 * 21 * 60 * 104 + 3
    lat, p, q = (_bf(torch.randn(n, device="cuda")) for _ in range(3))
    ref = lat.clone()
    pred = (q + (5.0 * (p - q)))  # bf16 ops round after each step, like the reference's tensor expression
    ref = ref + pred * torch.tensor(-0.0123)
    fwb.cfg_euler_step_(lat, p, q, 5.0, -0.0123)
    torch.testing.assert_close(lat.float(), ref.float(), rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("mma_n", [128, 96, 64])
def test_bringup_pv_mma_narrower_than_the_v_tile(fwb, mma_n):
    """tcgen05.mma in the attention's PV configuration (A = P from TMEM, B = V MN-major in two 64-column SWIZZLE_128B boxes) with an
    instruction N narrower than the loaded tile: the first mma_n columns of D must be exact.  N = 96 is the native head_dim-96 PV."""
    import ctypes as C
    torch.manual_seed(3)
    K, N = 128, 128
    A = _bf(torch.randn(128, K, device="cuda"))
    B = _bf(torch.randn(K, N, device="cuda"))             # [K][N]: MN-major operand, N contiguous
    D = torch.zeros(128, N, device="cuda", dtype=torch.float32)
    rc = fwb.lib.fwb_bringup_mma_pv_n(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mma_n, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, fwb.lib.fwb_last_error()
    torch.cuda.synchronize()
    ref = A.float() @ B.float()
    torch.testing.assert_close(D[:, :mma_n], ref[:, :mma_n], rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("variant", [1, 2])
def test_attention_head_dim_96_native_pv_width(fwb, variant):
    """fwb_attn_set_pv_n96: head_dim 96 with PV MMAs of N = 96 (no work on the zero-padded columns) == the padded N = 128 result."""
    torch.manual_seed(9)
    B, H, Lq, Lk, D = 1, 12, 1560, 1565, 96
    q, k, v = (_bf(torch.randn(B, L, H, D, device="cuda")) for L in (Lq, Lk, Lk))
    try:
        fwb.lib.fwb_attn_set_variant(variant)
        fwb.lib.fwb_attn_set_pv_n96(0)
        ref = fwb.attention(q, k, v)
        fwb.lib.fwb_attn_set_pv_n96(1)
        out = fwb.attention(q, k, v)
        torch.cuda.synchronize()
    finally:
        fwb.lib.fwb_attn_set_pv_n96(1)          # the default
        fwb.lib.fwb_attn_set_variant(0)
    assert torch.equal(out, ref)
    torch.testing.assert_close(out.float(), _attn_ref(q, k, v), rtol=2e-2, atol=6e-3)


@pytest.mark.parametrize("B,H,Lq,Lk,D", [(1, 4, 1024, 2777, 128), (1, 12, 2048, 2565, 96), (2, 3, 512, 3000, 128), (1, 40, 4096, 8190, 128)])
def test_attention_multicast_pairs_bit_identical(fwb, B, H, Lq, Lk, D):
    """fwb_attn_set_multicast: the aliased kernel as clusters of two CTAs (adjacent query blocks of one head) that share every K/V tile
    through TMA multicast and release a stage only when both have consumed it.  Same arithmetic in the same order per CTA: the output must
    equal the non-cluster kernel's bit for bit, in the bf16 and in the split-KV (fp32 partial + lse) modes."""
    torch.manual_seed(Lq + Lk)
    q, k, v = (_bf(torch.randn(B, L, H, D, device="cuda")) for L in (Lq, Lk, Lk))
    part = torch.empty(2, B, Lq, H, D, device="cuda")
    lse = torch.empty(2, B, H, Lq, device="cuda")
    try:
        fwb.lib.fwb_attn_set_variant(1)
        fwb.lib.fwb_attn_set_tail_split(0)
        fwb.lib.fwb_attn_set_multicast(0)
        ref = fwb.attention(q, k, v)
        fwb.attention_partial(q, k, v, part[0], lse[0])
        fwb.lib.fwb_attn_set_multicast(1)
        out = fwb.attention(q, k, v)
        fwb.attention_partial(q, k, v, part[1], lse[1])
        torch.cuda.synchronize()
    finally:
        fwb.lib.fwb_attn_set_multicast(1)       # the default
        fwb.lib.fwb_attn_set_tail_split(1)
        fwb.lib.fwb_attn_set_variant(0)
    assert torch.equal(out, ref)
    assert torch.equal(part[1], part[0]) and torch.equal(lse[1], lse[0])
    torch.testing.assert_close(out.float(), _attn_ref(q, k, v), rtol=2e-2, atol=6e-3)


@pytest.mark.parametrize("B,H,Lq,Lk,D,acc", [(1, 40, 2000, 512, 128, 0), (1, 40, 2000, 257, 128, 1), (21, 16, 1565, 1565, 64, 0), (1, 12, 700, 1565, 96, 0),
                                             (2, 3, 129, 1, 128, 0), (1, 2, 128, 2048, 64, 0)])
def test_attention_short_kv_configuration_bit_identical(fwb, B, H, Lq, Lk, D, acc):
    """fwb_attn_set_short_kv_max: with few keys the default policy runs the aliased kernel with one 128-row Q tile per CTA and two CTAs
    per SM (prologue / epilogue overlap).  A row's arithmetic does not depend on how rows are grouped into CTAs: the output (bf16,
    accumulate mode, fp32 partial + lse) must equal the two-tile configuration's bit for bit."""
    torch.manual_seed(Lq + Lk + D)
    q, k, v = (_bf(torch.randn(B, L, H, D, device="cuda")) for L in (Lq, Lk, Lk))
    base = _bf(torch.randn(B, Lq, H, D, device="cuda"))
    part = torch.empty(2, B, Lq, H, D, device="cuda")
    lse = torch.empty(2, B, H, Lq, device="cuda")
    try:
        fwb.lib.fwb_attn_set_short_kv_max(0)
        fwb.lib.fwb_attn_set_variant(1)
        fwb.lib.fwb_attn_set_tail_split(0)       # the two-tile schedule may split a part-filled wave along the keys (fp32 re-association)
        ref = fwb.attention(q, k, v, out=base.clone(), accumulate=bool(acc))
        fwb.attention_partial(q, k, v, part[0], lse[0])
        fwb.lib.fwb_attn_set_variant(0)
        fwb.lib.fwb_attn_set_short_kv_max(4096)
        out = fwb.attention(q, k, v, out=base.clone(), accumulate=bool(acc))
        fwb.attention_partial(q, k, v, part[1], lse[1])
        torch.cuda.synchronize()
    finally:
        fwb.lib.fwb_attn_set_short_kv_max(2048)
        fwb.lib.fwb_attn_set_tail_split(1)
        fwb.lib.fwb_attn_set_variant(0)
    assert torch.equal(out, ref)
    assert torch.equal(part[1], part[0]) and torch.equal(lse[1], lse[0])
    if not acc:
        torch.testing.assert_close(out.float(), _attn_ref(q, k, v), rtol=2e-2, atol=6e-3)
