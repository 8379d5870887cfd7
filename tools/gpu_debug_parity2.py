import sys, traceback
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / "fantasy-world_b200", ROOT / "tests", ROOT):
    sys.path.insert(0, str(p))
import torch
import fwb200
import test_gpu_parity as T
from _common import gold, rel_err
from fwb200.synth import build_fusion_model, synth_inputs

g = gold("joint_forward.pt")


def build():
    m = build_fusion_model(num_dit_layers=2, start_index=1, device="cuda", seed=0, heads=True, gen_device="cpu")
    m.vggt.depth_head.intermediate_layer_idx = g["head_layer_idx"]
    m.vggt.point_head.intermediate_layer_idx = g["head_layer_idx"]
    return m


def stages(model, tag):
    f, h, w = g["grid"]
    taps = {}
    hk = [model.pipe.dit.blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_pcb", o.clone())),
          model.vggt.aggregator.frame_blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_frame", o.clone())),
          model.vggt.aggregator.frame_blocks[0].register_forward_pre_hook(lambda m, a: taps.__setitem__("frame_in", a[0].clone())),
          model.IRGBlock[0].register_forward_hook(lambda m, i, o: taps.update(after_irg_x=o[0].clone(), after_irg_tokens=o[1].clone()))]
    inp = synth_inputs(f, h, w, device="cuda", seed=1024, text_len=g["text_len"])
    ts = torch.tensor([g["timestep"]], device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        out, _ = model.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], clip_feature=inp["clip_feature"],
                                     y=inp["y"], use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"])
    for x in hk:
        x.remove()
    line = [f"{tag}: out {rel_err(out.cpu(), g['out']):.4f}"]
    for k, v in g["taps"].items():
        line.append(f"{k} {rel_err(taps[k].float().cpu().reshape(v.shape), v):.4f}")
    fi = taps["frame_in"].float()
    line.append(f"frame_in absmean {fi.abs().mean().item():.4f} sum {fi.double().sum().item():.3f} dtype {taps['frame_in'].dtype}")
    print(" | ".join(line), flush=True)


print("=== A: IRG test function first, then stages ===")
mA = build()
try:
    T.test_irg_block_config1_vs_golden_and_oracle(mA)
    print("irg test passed")
except Exception:
    traceback.print_exc()
stages(mA, "A after irg test")
stages(mA, "A again")
print("=== B: fresh model, stages only ===")
mB = build()
stages(mB, "B first")
print("=== C: fresh model, the actual test functions in order ===")
mC = build()
for fn in (T.test_state_dict_roundtrip_with_reference_schema, T.test_irg_block_config1_vs_golden_and_oracle,
           T.test_joint_forward_with_heads_vs_golden, T.test_joint_forward_vs_bf16_oracle):
    try:
        fn(mC)
        print(fn.__name__, "PASSED", flush=True)
    except Exception as e:
        print(fn.__name__, "FAILED", str(e)[:300], flush=True)
