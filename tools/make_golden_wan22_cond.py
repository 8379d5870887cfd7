"""Golden for the Wan2.2 conditioning call (inference_wan22.py:345-353): the UNMODIFIED reference pipeline
FantasyWorld/diffsynth_wan22/pipelines/wan_video_new.py `WanVideoPipeline.__call__(..., return_condition=True)` on the CPU, with a reduced
umT5 (per-key synthetic weights), the Wan VAE (synthetic weights) and a stand-in tokenizer / DiT attribute bag; plus the reference's tiled
VAE ENCODE on its own and `preprocess_image` in bf16 (the pipeline dtype the scaling runs in).

    python tools/make_golden_wan22_cond.py      # build container -> tests/golden/wan22_condition.pt (~30 KB)
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

from make_golden_encoders import T5_CFG, FakeTokenizer  # noqa: E402

CALL = dict(prompt="a robot walks through  a\tquiet museum", negative_prompt="blurry low quality", seed=3, tiled=True, height=30, width=48,
            num_frames=6, tile_size=(2, 4), tile_stride=(1, 2), return_condition=True)     # 30 -> 32 rows, 6 -> 9 frames (shape check)


def images():
    g = torch.Generator().manual_seed(8)
    return (torch.randint(0, 256, (40, 60, 3), generator=g, dtype=torch.uint8), torch.randint(0, 256, (36, 52, 3), generator=g, dtype=torch.uint8))


def main():
    from PIL import Image
    from ref_shim import import_reference
    from fwb_synth import synth_init
    import_reference()
    pl = importlib.import_module("FantasyWorld.diffsynth_wan22.pipelines.wan_video_new")
    te = importlib.import_module("FantasyWorld.diffsynth_wan22.models.wan_video_text_encoder")
    vm = importlib.import_module("FantasyWorld.diffsynth_wan22.models.wan_video_vae")

    pipe = pl.WanVideoPipeline(device="cpu", torch_dtype=torch.float32)
    torch.manual_seed(0)
    pipe.text_encoder = synth_init(te.WanTextEncoder(**T5_CFG), seed=0, gen_device="cpu").eval()
    pipe.prompter.fetch_models(pipe.text_encoder)
    pipe.prompter.tokenizer = FakeTokenizer(24, T5_CFG["vocab"])
    wrap = nn.Module()
    wrap.vae = vm.WanVideoVAE(z_dim=16)
    wrap.vae.model.requires_grad_(True)
    synth_init(wrap, seed=0, gen_device="cpu")
    pipe.vae = wrap.vae.eval()
    pipe.height_division_factor = pipe.width_division_factor = pipe.vae.upsampling_factor * 2       # as from_pretrained does (:402-404)
    pipe.dit = types.SimpleNamespace(require_vae_embedding=True, require_clip_embedding=False, fuse_vae_embedding_in_latents=False,
                                     has_image_pos_emb=False, has_image_input=True, in_dim=36, control_adapter=None)
    a, b = (Image.fromarray(t.numpy()) for t in images())
    out = {"call": CALL, "t5_cfg": T5_CFG}
    with torch.no_grad():
        for tag, end in (("first", None), ("first_last", b)):
            shared, posi, nega = pipe(input_image=a, end_image=end, **CALL)
            out[tag] = {"y": shared["y"], "latents": shared["latents"], "noise": shared["noise"], "context_pos": posi["context"],
                        "context_neg": nega["context"], "height": shared["height"], "width": shared["width"],
                        "num_frames": shared["num_frames"], "timesteps": pipe.scheduler.timesteps.clone()}
        # the tiled encoder on its own, other tile geometry, and the untiled result for contrast
        g = torch.Generator().manual_seed(9)
        clip = (torch.rand(3, 5, 32, 48, generator=g) * 2 - 1)
        out["clip"] = clip
        out["enc_tiled"] = pipe.vae.encode([clip], device="cpu", tiled=True, tile_size=(3, 4), tile_stride=(2, 2))
        out["enc_single"] = pipe.vae.encode([clip], device="cpu", tiled=False)
    bf = pl.WanVideoPipeline(device="cpu", torch_dtype=torch.bfloat16)
    out["preprocess_bf16"] = bf.preprocess_image(a.resize((48, 32)))
    out["noise_bf16"] = bf.generate_noise((1, 16, 3, 4, 6), seed=5)
    out["pil_a"], out["pil_b"] = images()
    path = ROOT / "tests" / "golden" / "wan22_condition.pt"
    torch.save(out, path)
    print("wrote", path, path.stat().st_size, "bytes")
    for tag in ("first", "first_last"):
        r = out[tag]
        print(tag, tuple(r["y"].shape), tuple(r["context_pos"].shape), r["height"], r["width"], r["num_frames"], tuple(r["noise"].shape))
    print("tiled vs single encode rel diff", float((out["enc_tiled"] - out["enc_single"]).norm() / out["enc_single"].norm()))


if __name__ == "__main__":
    main()
