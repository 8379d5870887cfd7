"""Golden for the Wan2.2 reward-LoRA merge: the UNMODIFIED reference `load_lora`
(FantasyWorld/fusion/model_wan22.py:18-118) applied to a small WanModel, with LoRA files in the key styles it accepts
(kohya `lora_unet_<path>.lora_up/down.weight` + `.alpha`, PEFT `<path>.lora_A/B[.default].weight`, 1x1-conv factors).

    python tools/make_golden_lora.py       # build container -> tests/golden/lora_merge.pt (LoRA factors + merged weights, ~100 kB)

Also records which steps of the 50-step schedule take the high-noise expert under inference_wan22.py:229-240.
"""
from __future__ import annotations

import sys
import tempfile
import types
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

CFG = dict(dim=64, in_dim=36, ffn_dim=128, out_dim=16, text_dim=32, freq_dim=32, eps=1e-6, patch_size=(1, 2, 2), num_heads=2,
           num_layers=2, has_image_input=False)


def lora_files(seed=7, r=4):
    """Two LoRA state dicts (kohya style with alpha; PEFT style) over a few DiT layers."""
    g = torch.Generator().manual_seed(seed)

    def rn(*s):
        return torch.randn(*s, generator=g) * 0.2

    kohya = {}
    for path, (o, i) in {"blocks_0_self_attn_q": (64, 64), "blocks_1_cross_attn_k": (64, 64), "blocks_1_ffn_0": (128, 64)}.items():
        kohya[f"lora_unet_{path}.lora_up.weight"] = rn(o, r)
        kohya[f"lora_unet_{path}.lora_down.weight"] = rn(r, i)
        kohya[f"lora_unet_{path}.alpha"] = torch.tensor(2.0)
    peft, peft_default = {}, {}
    for path, (o, i) in {"blocks.0.self_attn.o": (64, 64), "blocks.0.ffn.2": (64, 128), "blocks.1.cross_attn.q": (64, 64)}.items():
        a, b = rn(r, i), rn(o, r)
        peft[f"{path}.lora_A.weight"], peft[f"{path}.lora_B.weight"] = a, b
        # the `.default` adapter-name variant: the reference strips 21 of the 22 characters of `_lora_A_default_weight`
        # (model_wan22.py:33-36), fails to resolve the layer and SKIPS it — recorded as such
        peft_default[f"{path}.lora_A.default.weight"], peft_default[f"{path}.lora_B.default.weight"] = a, b
    return {"kohya": kohya, "peft": peft, "peft_default": peft_default}


def main():
    from ref_shim import import_reference
    from fwb_synth import synth_init
    import importlib
    from safetensors.torch import save_file
    import_reference()
    ref22 = importlib.import_module("FantasyWorld.fusion.model_wan22")
    dit22 = importlib.import_module("FantasyWorld.diffsynth_wan22.models.wan_video_dit")
    sched_mod = importlib.import_module("FantasyWorld.diffsynth_wan22.schedulers.flow_match")
    out = {"cfg": CFG, "loras": lora_files(), "merged": {}}
    for style, sd in out["loras"].items():
        torch.manual_seed(0)
        model = dit22.WanModel(**CFG)
        wrap = torch.nn.Module()
        wrap.dit = model
        synth_init(wrap, seed=0, gen_device="cpu")
        model.to(torch.bfloat16)
        pipe = types.SimpleNamespace(device="cpu", torch_dtype=torch.bfloat16, dit=model)
        with tempfile.TemporaryDirectory() as d:
            f = Path(d) / "lora.safetensors"
            save_file({k: v.contiguous() for k, v in sd.items()}, str(f))
            ref22.load_lora(pipe, str(f), 0.55, "dit")
        keep = ("blocks.0.self_attn.q.", "blocks.1.cross_attn.k.", "blocks.1.ffn.0.", "blocks.0.self_attn.o.", "blocks.0.ffn.2.",
                "blocks.1.cross_attn.q.", "blocks.1.self_attn.v.")      # the touched layers of both styles + one untouched
        out["merged"][style] = {k: v.clone() for k, v in model.state_dict().items() if k.endswith("weight") and k.startswith(keep)}
    sched = sched_mod.FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    sched.set_timesteps(50)
    out["high_noise_steps"] = [bool(t.unsqueeze(0).to(torch.bfloat16).item() > 900.0) for t in sched.timesteps]
    path = ROOT / "tests" / "golden" / "lora_merge.pt"
    torch.save(out, path)
    print("wrote", path, path.stat().st_size, "bytes; high-noise steps:", sum(out["high_noise_steps"]))


if __name__ == "__main__":
    main()
