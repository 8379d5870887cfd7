"""Golden for the `uncond=True` branch of IRGBlock.forward (fusion/layer/block.py:70-72: the bidirectional adapter is skipped),
generated from the UNMODIFIED reference on CPU in fp32 with the same synthetic weights and inputs as irg_block_c1
(tools/make_golden.py).  Build container only.

    python tools/make_golden_uncond.py  ->  tests/golden/irg_block_c1_uncond.pt
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
GOLD = ROOT / "tests" / "golden"
TEXT_LEN = 64


def main():
    from ref_shim import build_reference_fusion
    from fwb_synth import synth_init

    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    model, ns = build_reference_fusion(num_dit_layers=2, start_index=1, heads=False)
    synth_init(model, seed=0, gen_device="cpu")
    g = torch.Generator().manual_seed(1024)
    f, h, w = 1, 4, 4
    L = f * h * w
    x_dit = torch.randn(1, L, 5120, generator=g)
    x_agg = torch.randn(f, 5 + h * w, 1024, generator=g)
    context = torch.randn(1, 257 + TEXT_LEN, 5120, generator=g)
    t_mod = torch.randn(1, 6, 5120, generator=g) * 0.1
    e0 = torch.randn(1, 6, 1024, generator=g) * 0.1
    plucker = torch.randn(1, L, 2048, generator=g)
    dit = ns.dit
    freqs3 = dit.precompute_freqs_cis_3d(128)
    freqs = torch.cat([freqs3[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), freqs3[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                       freqs3[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(L, 1, -1)
    fb = model.freqs_bicross
    freqs_bi_dit = torch.cat([fb[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), fb[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                              fb[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(L, 1, -1)
    freqs_bi_agg = dit.build_freqs_3d_with_extra_cis(fb, f, h, w, n_extra=5)
    pos = model.vggt.aggregator.position_getter(f, h, w, device="cpu") + 1
    pos = torch.cat([torch.zeros(f, 5, 2, dtype=pos.dtype), pos], dim=1)
    xd, xa, _ = model.IRGBlock[0](x_dit=x_dit, x_agg=x_agg, context=context, t_mod=t_mod, freqs=freqs, freqs_dit=freqs_bi_dit,
                                  freqs_agg=freqs_bi_agg, pos=pos, e0=e0, uncond=True, plucker_fea=plucker,
                                  plucker_context_lens=torch.ones(1, dtype=torch.long))
    torch.save({"x_dit_out": xd.clone(), "x_agg_out": xa.clone(), "seed": 1024, "text_len": TEXT_LEN}, GOLD / "irg_block_c1_uncond.pt")
    print("irg_block_c1_uncond done", xd.abs().mean().item(), xa.abs().mean().item())


if __name__ == "__main__":
    main()
