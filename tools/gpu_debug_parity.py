"""Stage-by-stage parity of the reduced joint_forward on the GPU against the reference golden taps (debug aid)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
import contextlib
import torch
import fwb200
from _common import gold, rel_err
from fwb200.synth import build_fusion_model, synth_inputs

g = gold("joint_forward.pt")
heads = "heads" in sys.argv
model = build_fusion_model(num_dit_layers=2, start_index=1, device="cuda", seed=0, heads=heads, gen_device="cpu")
if heads:
    model.vggt.depth_head.intermediate_layer_idx = g["head_layer_idx"]
    model.vggt.point_head.intermediate_layer_idx = g["head_layer_idx"]
f, h, w = g["grid"]
taps = {}
model.pipe.dit.blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_pcb", o.clone()))
model.vggt.aggregator.frame_blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_frame", o.clone()))
model.IRGBlock[0].register_forward_hook(lambda m, i, o: taps.update(after_irg_x=o[0].clone(), after_irg_tokens=o[1].clone()))


def run(tag, autocast=False, pred=False, lens=False):
    inp = synth_inputs(f, h, w, device="cuda", seed=1024, text_len=g["text_len"])
    ts = torch.tensor([g["timestep"]], device="cuda", dtype=torch.bfloat16)
    ln = None
    if lens:
        ln = torch.ones(f, dtype=torch.long, device="cuda")
        ln[1:] = 4
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with torch.no_grad(), ctx:
        out, _ = model.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], clip_feature=inp["clip_feature"],
                                     y=inp["y"], use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"],
                                     plucker_context_lens=ln, return_prediction=pred and heads)
    line = [f"{tag}: out rel {rel_err(out.cpu(), g['out']):.4f}"]
    for k, v in g["taps"].items():
        t = taps[k].float().cpu().reshape(v.shape)
        line.append(f"{k} {rel_err(t, v):.4f}")
    print(" | ".join(line), flush=True)


run("plain")
if "irgfirst" in sys.argv:
    gi = gold("irg_block_c1.pt")
    gen = torch.Generator().manual_seed(gi["seed"])
    L = 16
    x_dit = torch.randn(1, L, 5120, generator=gen); x_agg = torch.randn(1, 21, 1024, generator=gen)
    context = torch.randn(1, 257 + gi["text_len"], 5120, generator=gen); t_mod = torch.randn(1, 6, 5120, generator=gen) * 0.1
    e0 = torch.randn(1, 6, 1024, generator=gen) * 0.1; plucker = torch.randn(1, L, 2048, generator=gen)
    fr, fd, fa = model.rope_tables(1, 4, 4, "cuda")
    pos = model.vggt.aggregator._positions(1, 4, 4, torch.device("cuda"))
    bf = torch.bfloat16
    with torch.no_grad():
        model.IRGBlock[0](x_dit=x_dit.cuda().to(bf), x_agg=x_agg.cuda().to(bf), context=context.cuda().to(bf), t_mod=t_mod.cuda().to(bf),
                          freqs=fr, freqs_dit=fd, freqs_agg=fa, pos=pos, e0=e0.cuda(), uncond=False, plucker_fea=plucker.cuda().to(bf),
                          plucker_context_lens=torch.ones(1, dtype=torch.long))
    run("after irg call")
run("autocast", autocast=True)
run("autocast+lens", autocast=True, lens=True)
run("autocast+pred", autocast=True, pred=True, lens=True)
run("plain again")
