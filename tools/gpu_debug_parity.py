"""Stage-by-stage parity of the reduced joint_forward on the GPU against the reference golden taps (debug aid)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT))
import torch
import fwb200
from _common import gold, rel_err
from fwb200.synth import build_fusion_model, synth_inputs

g = gold("joint_forward.pt")
model = build_fusion_model(num_dit_layers=2, start_index=1, device="cuda", seed=0, heads=False, gen_device="cpu")
f, h, w = g["grid"]
taps = {}
model.pipe.dit.blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_pcb", o.clone()))
model.vggt.aggregator.frame_blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_frame", o.clone()))
model.IRGBlock[0].register_forward_hook(lambda m, i, o: taps.update(after_irg_x=o[0].clone(), after_irg_tokens=o[1].clone()))
for rep in range(3):
    inp = synth_inputs(f, h, w, device="cuda", seed=1024, text_len=g["text_len"])
    ts = torch.tensor([g["timestep"]], device="cuda", dtype=torch.bfloat16)
    try:
        with torch.no_grad():
            out, _ = model.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], clip_feature=inp["clip_feature"],
                                         y=inp["y"], use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"])
    except FloatingPointError as e:
        print("rep", rep, "FloatingPointError:", e)
        continue
    line = [f"rep {rep}: out rel {rel_err(out.cpu(), g['out']):.4f}"]
    for k, v in g["taps"].items():
        t = taps[k].float().cpu().reshape(v.shape)
        line.append(f"{k} {rel_err(t, v):.4f} nan={bool(torch.isnan(t).any())}")
    print(" | ".join(line))
