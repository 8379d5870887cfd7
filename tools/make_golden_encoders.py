"""Goldens for the conditioning front-end mirrors: the UNMODIFIED reference umT5 encoder
(FantasyWorld/diffsynth_wan21/models/wan_video_text_encoder.py), CLIP image tower (…/wan_video_image_encoder.py), prompter
(…/prompters/wan_prompter.py) and `WanVideoPipeline.encode_image` (…/pipelines/wan_video.py:218-276) on the CPU.

    python tools/make_golden_encoders.py     # build container -> tests/golden/encoders.pt (~0.3 MB)

Reduced towers (the full ones are 5.7 B / 0.63 B parameters) with the per-key synthetic weights of fwb_synth, so the test rebuilds
identical weights in the mirror; each forward is stored twice — fp32 (the yardstick) and the reference's own bf16 run (what the
pipeline executes; its distance to fp32 is the error budget of the CUDA mirror).  Full-size state_dict schemas are compared here
key by key against the mirrors (meta device) and pinned as sha256 digests for the boxes that have no /root/reference.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

T5_CFG = dict(vocab=64, dim=64, dim_attn=128, dim_ffn=128, num_heads=2, num_layers=2, num_buckets=32, shared_pos=False, dropout=0.1)
T5_SHARED_CFG = dict(T5_CFG, num_layers=1, shared_pos=True)
CLIP_CFG = dict(embed_dim=64, image_size=56, patch_size=14, vision_dim=160, vision_heads=2, vision_layers=3)


def schema_digest(sd) -> str:
    return hashlib.sha256(json.dumps(sorted((k, list(v.shape)) for k, v in sd.items())).encode()).hexdigest()


def t5_inputs():
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, T5_CFG["vocab"], (2, 24), generator=g)
    mask = torch.ones(2, 24, dtype=torch.long)
    mask[0, 15:] = 0
    return ids, mask


def clip_inputs():
    g = torch.Generator().manual_seed(6)
    return (torch.rand(1, 3, 40, 72, generator=g) * 2 - 1), torch.randint(0, 256, (36, 52, 3), generator=g, dtype=torch.uint8)


class FakeTokenizer:
    """Stand-in for HuggingfaceTokenizer (no umT5 tokenizer files offline): whitespace words -> ids, right-padded."""

    def __init__(self, seq_len, vocab):
        self.seq_len, self.vocab = seq_len, vocab

    def __call__(self, sequence, return_mask=False, add_special_tokens=True):
        if isinstance(sequence, str):
            sequence = [sequence]
        ids = torch.zeros(len(sequence), self.seq_len, dtype=torch.long)
        mask = torch.zeros_like(ids)
        for i, s in enumerate(sequence):
            toks = [2 + sum(map(ord, w)) % (self.vocab - 2) for w in s.split()][: self.seq_len - 1] + [1]
            ids[i, :len(toks)] = torch.tensor(toks)
            mask[i, :len(toks)] = 1
        return (ids, mask) if return_mask else ids


def both(model: nn.Module, run):
    """run(model) in fp32, then the same module converted to bf16 (CPU bf16 kernels)."""
    with torch.no_grad():
        f = run(model, torch.float32)
        model.to(torch.bfloat16)
        b = run(model, torch.bfloat16)
        model.to(torch.float32)
    return f, b


def main():
    from ref_shim import import_reference
    from fwb_synth import synth_init
    import_reference()
    te = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder")
    ie = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_image_encoder")
    wp = importlib.import_module("FantasyWorld.diffsynth_wan21.prompters.wan_prompter")
    pl = importlib.import_module("FantasyWorld.diffsynth_wan21.pipelines.wan_video")
    vae_mod = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_vae")
    out = {"t5_cfg": T5_CFG, "t5_shared_cfg": T5_SHARED_CFG, "clip_cfg": CLIP_CFG}

    # ---- full-size schemas (meta) -----------------------------------------------------------------------------------------
    with torch.device("meta"):
        full_t5 = te.WanTextEncoder()
        clip_full = ie.XLMRobertaCLIP(**dict(embed_dim=1024, image_size=224, patch_size=14, vision_dim=1280, vision_mlp_ratio=4,
                                             vision_heads=16, vision_layers=32, vision_pool="token", activation="gelu"))
    full_ie = {"model." + k: v for k, v in clip_full.state_dict().items()}
    out["schema_sha"] = {"t5": schema_digest(full_t5.state_dict()), "clip": schema_digest(full_ie)}
    out["schema_len"] = {"t5": len(full_t5.state_dict()), "clip": len(full_ie)}

    # ---- umT5 (reduced) -----------------------------------------------------------------------------------------------------
    ids, mask = t5_inputs()
    for tag, cfg in (("t5", T5_CFG), ("t5_shared", T5_SHARED_CFG)):
        torch.manual_seed(0)
        m = synth_init(te.WanTextEncoder(**cfg), seed=0, gen_device="cpu").eval()
        out[tag + "_schema"] = {k: list(v.shape) for k, v in m.state_dict().items()}
        f, b = both(m, lambda mod, dt: mod(ids, mask))
        out[tag + "_fp32"], out[tag + "_bf16"] = f, b.float()
        if tag == "t5":
            # prompter: tokenise -> encode -> zero the padding
            pr = wp.WanPrompter(tokenizer_path=None, text_len=24)
            pr.tokenizer = FakeTokenizer(24, cfg["vocab"])
            pr.fetch_models(m)
            with torch.no_grad():
                out["prompt_emb"] = pr.encode_prompt(["a robot walks through  a\tquiet museum", "sunrise"], device="cpu")
    rel = torch.arange(-700, 701)[None, :]
    out["buckets_bidir"] = te.T5RelativeEmbedding(32, 2, bidirectional=True)._relative_position_bucket(rel.clone())
    out["buckets_unidir"] = te.T5RelativeEmbedding(32, 2, bidirectional=False)._relative_position_bucket(rel.clone())

    # ---- text cleaning helpers (pure python) ----------------------------------------------------------------------------------
    samples = ["  Hello,\n  WORLD_of   <b>tags</b> &amp;amp; more!  ", "snake_case_and-dashes; (parens)", "keep <|sep|> this_marker <|sep|> ok?"]
    out["clean"] = {"samples": samples, "whitespace": [wp.whitespace_clean(s) for s in samples],
                    "canonicalize": [wp.canonicalize(s) for s in samples],
                    "canonicalize_keep": [wp.canonicalize(s, keep_punctuation_exact_string="<|sep|>") for s in samples]}

    # ---- CLIP image tower (reduced) ------------------------------------------------------------------------------------------
    torch.manual_seed(0)
    clip, transforms = ie.clip_xlm_roberta_vit_h_14(pretrained=False, return_transforms=True, return_tokenizer=False,
                                                    dtype=torch.float32, device="cpu", **CLIP_CFG)
    enc = ie.WanImageEncoder.__new__(ie.WanImageEncoder)      # its __init__ hard-codes ViT-H/14; the methods are what we pin
    nn.Module.__init__(enc)
    enc.model, enc.transforms = clip, transforms
    synth_init(enc, seed=0, gen_device="cpu").eval()
    out["clip_schema"] = {k: list(v.shape) for k, v in enc.state_dict().items()}
    img, pil_arr = clip_inputs()
    f, b = both(enc, lambda mod, dt: mod.encode_image([img.clone()]))
    out["clip_fp32"], out["clip_bf16"] = f, b.float()
    x224 = torch.nn.functional.interpolate(img, size=(56, 56), mode="bicubic", align_corners=False)
    with torch.no_grad():
        out["clip_all_blocks_fp32"] = clip.visual(x224)
        out["clip_interp_fp32"] = clip.visual(torch.nn.functional.interpolate(img, size=(70, 70), mode="bicubic", align_corners=False),
                                              interpolation=True)
        out["pos_interp"] = ie.pos_interpolate(clip.visual.pos_embedding.detach(), 26)

    # ---- pipeline.encode_image (first-frame conditioning: CLIP tokens + mask / VAE-latent `y`) --------------------------------
    from PIL import Image
    pipe = pl.WanVideoPipeline(device="cpu", torch_dtype=torch.float32)
    pipe.image_encoder = enc
    wrap = nn.Module()
    wrap.vae = vae_mod.WanVideoVAE(z_dim=16)
    wrap.vae.model.requires_grad_(True)
    synth_init(wrap, seed=0, gen_device="cpu")
    pipe.vae = wrap.vae.eval()
    pipe.dit = types.SimpleNamespace(has_image_pos_emb=False)
    pil = Image.fromarray(pil_arr.numpy())
    with torch.no_grad():
        r = pipe.encode_image(pil, None, 5, 32, 48)
        r2 = pipe.encode_image(pil, pil.transpose(Image.FLIP_LEFT_RIGHT), 5, 32, 48)
    out["pipe_y"], out["pipe_clip"] = r["y"], r["clip_feature"]
    out["pipe_y_end"] = r2["y"]
    out["pil"] = pil_arr

    path = ROOT / "tests" / "golden" / "encoders.pt"
    torch.save(out, path)
    print("wrote", path, path.stat().st_size, "bytes")
    for k in ("t5_fp32", "clip_fp32", "pipe_y", "pipe_clip"):
        print(k, tuple(out[k].shape))
    for tag in ("t5", "clip"):
        ref, got = out[tag + "_fp32"], out[tag + "_bf16"]
        print(tag, "reference bf16-vs-fp32 rel err", float((got - ref).norm() / ref.norm()))

    # ---- the mirrors must expose the same full-size schemas -----------------------------------------------------------------
    for name in [m for m in sys.modules if m == "FantasyWorld" or m.startswith("FantasyWorld.")]:
        del sys.modules[name]
    from ref_shim import REF_ROOT
    sys.path.remove(REF_ROOT)
    mt = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder")
    mi = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_image_encoder")
    assert "fantasy-world_b200" in mt.__file__
    with torch.device("meta"):
        ours_t5 = mt.WanTextEncoder()
    ours_ie = mi.WanImageEncoder(device="meta")
    assert schema_digest(ours_t5.state_dict()) == out["schema_sha"]["t5"], "T5 schema differs"
    assert schema_digest(ours_ie.state_dict()) == out["schema_sha"]["clip"], "CLIP schema differs"
    print("full-size schemas match:", out["schema_len"])


if __name__ == "__main__":
    main()
