"""Conditioning front-end mirrors: umT5 text encoder, CLIP ViT-H image tower, prompter, pipeline.encode_image
against goldens written by the UNMODIFIED reference on the CPU (tools/make_golden_encoders.py).

CPU suite: schemas (reduced key-by-key, full size by digest), bucket / interpolation / text-cleaning helpers bit-equal, and the
mirrors' HOST LOGIC run on the test-only torch shim of the four fwb200 ops (tests/_ops_torch_shim.py).
GPU suite (-m gpu): the same forwards on the real kernels, held to the reference's own bf16 error budget, plus full-size
(5.7 B / 0.63 B parameter) forwards compared with the fp32-accumulating shim on the same weights.
"""
import hashlib
import json

import pytest
import torch

from _common import gold, rel_err
from _ops_torch_shim import torch_ops

BF16 = torch.bfloat16


def _digest(sd):
    return hashlib.sha256(json.dumps(sorted((k, list(v.shape)) for k, v in sd.items())).encode()).hexdigest()


def _t5(cfg, device="cpu"):
    from FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder import WanTextEncoder
    from fwb_synth import synth_init
    torch.manual_seed(0)
    return synth_init(WanTextEncoder(**cfg), seed=0, gen_device="cpu").to(device=device, dtype=BF16).eval()


def _clip(cfg, device="cpu"):
    from FantasyWorld.diffsynth_wan21.models.wan_video_image_encoder import WanImageEncoder
    from fwb_synth import synth_init
    torch.manual_seed(0)
    return synth_init(WanImageEncoder(**cfg), seed=0, gen_device="cpu").to(device=device, dtype=BF16).eval()


def _t5_inputs(vocab):
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, vocab, (2, 24), generator=g)
    mask = torch.ones(2, 24, dtype=torch.long)
    mask[0, 15:] = 0
    return ids, mask


def _clip_inputs():
    g = torch.Generator().manual_seed(6)
    return (torch.rand(1, 3, 40, 72, generator=g) * 2 - 1), torch.randint(0, 256, (36, 52, 3), generator=g, dtype=torch.uint8)


class FakeTokenizer:
    """Same stand-in as tools/make_golden_encoders.py (no umT5 tokenizer files offline)."""

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


def _budget(g, tag, out, slack=1.5, floor=2e-3):
    """Error against the fp32 golden no worse than `slack` x the reference's own bf16 run (+ a small floor)."""
    ref = g[tag + "_fp32"]
    ours, theirs = rel_err(out, ref), rel_err(g[tag + "_bf16"], ref)
    assert ours <= slack * theirs + floor, f"{tag}: ours {ours:.3e} vs reference bf16 {theirs:.3e}"
    return ours, theirs


# ---------------------------------------------------------------------------------------------------------------------------
# schemas and pure host helpers
# ---------------------------------------------------------------------------------------------------------------------------
def test_full_size_schemas_match_reference_digests():
    from FantasyWorld.diffsynth_wan21.models.wan_video_image_encoder import WanImageEncoder
    from FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder import WanTextEncoder
    g = gold("encoders.pt")
    with torch.device("meta"):
        t5 = WanTextEncoder()
    clip = WanImageEncoder(device="meta")
    assert len(t5.state_dict()) == g["schema_len"]["t5"] == 242 and _digest(t5.state_dict()) == g["schema_sha"]["t5"]
    assert len(clip.state_dict()) == g["schema_len"]["clip"] == 393 and _digest(clip.state_dict()) == g["schema_sha"]["clip"]
    assert sum(p.numel() for p in t5.parameters()) == 5_680_910_336          # umT5-XXL encoder
    # the released CLIP checkpoint also carries the text tower: the converter drops it and prefixes the rest
    sd = {"visual.pre_norm.weight": 1, "textual.token_embedding.weight": 2, "log_scale": 3}
    assert WanImageEncoder.state_dict_converter().from_civitai(sd) == {"model.visual.pre_norm.weight": 1, "model.log_scale": 3}


def test_reduced_schemas_match_reference():
    g = gold("encoders.pt")
    assert {k: list(v.shape) for k, v in _t5(g["t5_cfg"]).state_dict().items()} == g["t5_schema"]
    assert {k: list(v.shape) for k, v in _t5(g["t5_shared_cfg"]).state_dict().items()} == g["t5_shared_schema"]
    assert {k: list(v.shape) for k, v in _clip(g["clip_cfg"]).state_dict().items()} == g["clip_schema"]


def test_relative_position_buckets_bit_equal():
    from FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder import T5RelativeEmbedding
    g = gold("encoders.pt")
    rel = torch.arange(-700, 701)[None, :]
    assert torch.equal(T5RelativeEmbedding(32, 2, bidirectional=True)._relative_position_bucket(rel.clone()), g["buckets_bidir"])
    assert torch.equal(T5RelativeEmbedding(32, 2, bidirectional=False)._relative_position_bucket(rel.clone()), g["buckets_unidir"])
    e = T5RelativeEmbedding(32, 2, bidirectional=True)
    assert e(5, 7).shape == (1, 2, 5, 7)


def test_text_cleaning_matches_reference():
    from FantasyWorld.diffsynth_wan21.prompters import wan_prompter as wp
    c = gold("encoders.pt")["clean"]
    assert [wp.whitespace_clean(s) for s in c["samples"]] == c["whitespace"]
    assert [wp.canonicalize(s) for s in c["samples"]] == c["canonicalize"]
    assert [wp.canonicalize(s, keep_punctuation_exact_string="<|sep|>") for s in c["samples"]] == c["canonicalize_keep"]


def test_pos_interpolate_matches_reference():
    from FantasyWorld.diffsynth_wan21.models.wan_video_image_encoder import pos_interpolate
    g = gold("encoders.pt")
    enc = _clip(g["clip_cfg"]).float()
    pos = enc.model.visual.pos_embedding.detach()
    assert pos_interpolate(pos, pos.shape[1]) is pos
    out = pos_interpolate(pos, 26)
    # weights were rounded to bf16 by _clip(); the golden used fp32 weights
    assert out.shape == g["pos_interp"].shape and rel_err(out, g["pos_interp"]) < 4e-3


def test_model_manager_detects_side_checkpoints():
    from FantasyWorld.diffsynth_wan21.models.model_manager import ModelManager, detect_pth
    from FantasyWorld.diffsynth_wan21.models.wan_video_vae import WanVideoVAE
    assert detect_pth(["token_embedding.weight", "blocks.0.attn.q.weight"]) == "wan_video_text_encoder"
    assert detect_pth(["visual.patch_embedding.weight", "textual.x"]) == "wan_video_image_encoder"
    assert detect_pth(["encoder.conv1.weight", "decoder.conv1.weight"]) == "wan_video_vae"
    assert detect_pth(["blocks.0.self_attn.q.weight"]) is None
    # a VAE checkpoint in the released layout (bare VideoVAE_ keys under 'model_state') loads strictly through the manager
    vae = WanVideoVAE(z_dim=16)
    ckpt = {"model_state": {k[len("model."):]: v for k, v in vae.state_dict().items()}}
    mm = ModelManager(torch_dtype=torch.float32, device="cpu")
    assert mm.load_state_dict_model(ckpt, path="/x/Wan2.1_VAE.pth") == "wan_video_vae"
    model, path = mm.fetch_model("wan_video_vae", require_model_path=True)
    assert path == "/x/Wan2.1_VAE.pth" and isinstance(model, WanVideoVAE)
    assert mm.fetch_model("wan_video_text_encoder", require_model_path=True) is None


def test_pipeline_fetches_side_models_from_checkpoints(tmp_path):
    """inference_wan21.py:183-188 hands the DiT shards plus three `.pth` files to the model manager; the pipeline must come back with the
    text encoder, the image encoder and the VAE attached (their keys are part of the fusion model's state_dict: `pipe.text_encoder.*`
    ... — the CLI asserts that the released checkpoint has no unexpected keys).  Reduced towers, files in the released layouts."""
    from FantasyWorld.diffsynth_wan21.models.model_manager import ModelManager
    from FantasyWorld.diffsynth_wan21.models.wan_video_vae import WanVideoVAE
    from FantasyWorld.diffsynth_wan21.pipelines.wan_video import WanVideoPipeline
    g = gold("encoders.pt")
    t5, clip, vae = _t5(g["t5_cfg"]), _clip(g["clip_cfg"]), WanVideoVAE(z_dim=16)
    t5_path, clip_path, vae_path = (tmp_path / n for n in ("models_t5_umt5-xxl-enc-bf16.pth", "models_clip.pth", "Wan2.1_VAE.pth"))
    torch.save(t5.state_dict(), t5_path)
    clip_sd = {k[len("model."):]: v for k, v in clip.state_dict().items()}
    clip_sd["textual.token_embedding.weight"] = torch.zeros(4, 4)                  # the released file also holds the text tower
    torch.save(clip_sd, clip_path)
    torch.save({k[len("model."):]: v for k, v in vae.state_dict().items()}, vae_path)
    mm = ModelManager(torch_dtype=BF16, device="cpu", dit_config=dict(dim=64, in_dim=36, ffn_dim=128, out_dim=16, text_dim=64, freq_dim=32,
                                                                        eps=1e-6, patch_size=(1, 2, 2), num_heads=2, num_layers=1,
                                                                        has_image_input=True),
                      side_configs={"wan_video_text_encoder": g["t5_cfg"], "wan_video_image_encoder": g["clip_cfg"]})
    mm.load_models([[], str(vae_path), str(clip_path), str(t5_path)])
    pipe = WanVideoPipeline.from_model_manager(mm, device="cpu")
    assert pipe.dit is not None and pipe.vae is not None and pipe.image_encoder is not None and pipe.text_encoder is not None
    assert pipe.prompter.text_encoder is pipe.text_encoder and pipe.prompter.tokenizer is None      # no google/umt5-xxl dir next to the file
    for ours, src in ((pipe.text_encoder, t5), (pipe.image_encoder, clip)):
        a, b = ours.state_dict(), src.state_dict()
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    keys = {k for k, _ in pipe.named_parameters()}
    assert "text_encoder.blocks.0.attn.q.weight" in keys and "image_encoder.model.visual.pre_norm.weight" in keys
    assert any(k.startswith("vae.model.") for k in keys)


def test_encoders_refuse_to_run_without_the_device():
    """No CPU fallback: without a B200 the mirrors raise instead of computing something else."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    g = gold("encoders.pt")
    ids, mask = _t5_inputs(g["t5_cfg"]["vocab"])
    with pytest.raises(RuntimeError, match="B200"):
        _t5(g["t5_cfg"])(ids, mask)
    with pytest.raises(RuntimeError, match="B200"):
        _clip(g["clip_cfg"]).encode_image([_clip_inputs()[0]])
    from FantasyWorld.diffsynth_wan21.pipelines.wan_video import WanVideoPipeline
    pipe = WanVideoPipeline(device="cpu")
    with pytest.raises(RuntimeError, match="text encoder"):
        pipe.encode_prompt("x")
    with pytest.raises(RuntimeError, match="image encoder"):
        pipe.encode_image(None, None, 5, 32, 48)


# ---------------------------------------------------------------------------------------------------------------------------
# forwards: shared by the CPU (torch shim) and GPU (real kernels) variants
# ---------------------------------------------------------------------------------------------------------------------------
def _check_t5(device):
    g = gold("encoders.pt")
    ids, mask = _t5_inputs(g["t5_cfg"]["vocab"])
    res = {}
    for tag in ("t5", "t5_shared"):
        m = _t5(g[tag + "_cfg"], device)
        out = m(ids.to(device), mask.to(device))
        assert out.dtype == BF16 and out.shape == g[tag + "_fp32"].shape
        res[tag] = _budget(g, tag, out.cpu())
    # prompter: tokenise -> encode -> padding zeroed from the shortest prompt's length on (reference behaviour)
    from FantasyWorld.diffsynth_wan21.prompters import WanPrompter
    pr = WanPrompter(tokenizer_path=None, text_len=24)
    pr.tokenizer = FakeTokenizer(24, g["t5_cfg"]["vocab"])
    pr.fetch_models(_t5(g["t5_cfg"], device))
    emb = pr.encode_prompt(["a robot walks through  a\tquiet museum", "sunrise"], device=device).cpu()
    ref = g["prompt_emb"]
    assert torch.equal((emb == 0).all(-1), (ref == 0).all(-1)) and bool((emb[:, 2:] == 0).all()) and rel_err(emb, ref) < 0.12
    return res


def _check_clip(device):
    g = gold("encoders.pt")
    enc = _clip(g["clip_cfg"], device)
    img, _ = _clip_inputs()
    out = enc.encode_image([img.to(device)])
    assert out.dtype == BF16 and out.shape == (1, 17, 160)
    res = {"clip": _budget(g, "clip", out.cpu())}
    x = torch.nn.functional.interpolate(img, size=(56, 56), mode="bicubic", align_corners=False).to(device, BF16)
    allb = enc.model.visual(x).cpu()
    assert rel_err(allb, g["clip_all_blocks_fp32"]) < 1.5 * res["clip"][1] + 4e-3
    x70 = torch.nn.functional.interpolate(img, size=(70, 70), mode="bicubic", align_corners=False).to(device, BF16)
    itp = enc.model.visual(x70, interpolation=True).cpu()
    assert itp.shape == g["clip_interp_fp32"].shape == (1, 26, 160) and rel_err(itp, g["clip_interp_fp32"]) < 1.5 * res["clip"][1] + 4e-3
    return res


def _check_pipeline_encode_image(device):
    from PIL import Image
    from FantasyWorld.diffsynth_wan21.pipelines.wan_video import WanVideoPipeline
    from fwb_synth import synth_init
    import types
    g = gold("encoders.pt")
    pipe = WanVideoPipeline(device=device, torch_dtype=torch.float32)
    pipe.image_encoder = _clip(g["clip_cfg"], device)
    vae = pipe.enable_vae(z_dim=16, device=device, dtype=torch.float32)
    wrap = torch.nn.Module()
    wrap.vae = vae
    vae.model.requires_grad_(True)
    synth_init(wrap, seed=0, gen_device="cpu")
    vae.model.requires_grad_(False)
    pipe.dit = types.SimpleNamespace(has_image_pos_emb=False)
    pil = Image.fromarray(g["pil"].numpy())
    r = pipe.encode_image(pil, None, 5, 32, 48)
    r2 = pipe.encode_image(pil, pil.transpose(Image.FLIP_LEFT_RIGHT), 5, 32, 48)
    y, y2 = r["y"].cpu(), r2["y"].cpu()
    assert y.shape == g["pipe_y"].shape == (1, 20, 2, 4, 6)
    assert torch.equal(y[:, :4], g["pipe_y"][:, :4]) and torch.equal(y2[:, :4], g["pipe_y_end"][:, :4])      # mask channels: exact
    tol = 2e-5 if device == "cpu" else 2e-3         # fp32 VAE: cuDNN may use TF32 convolutions, like the reference on the GPU
    assert rel_err(y[:, 4:], g["pipe_y"][:, 4:]) < tol and rel_err(y2[:, 4:], g["pipe_y_end"][:, 4:]) < tol
    assert r["clip_feature"].dtype == torch.float32 and rel_err(r["clip_feature"].cpu(), g["pipe_clip"]) < 0.03


def test_t5_host_logic_on_torch_shim():
    with torch_ops():
        res = _check_t5("cpu")
    print("t5 (shim) rel err vs fp32 golden, reference bf16:", res)


def test_clip_host_logic_on_torch_shim():
    with torch_ops():
        res = _check_clip("cpu")
    print("clip (shim):", res)


def test_pipeline_encode_image_host_logic_on_torch_shim():
    with torch_ops():
        _check_pipeline_encode_image("cpu")


@pytest.mark.gpu
def test_t5_matches_reference_on_cuda():
    print("t5 (fwb200 kernels):", _check_t5("cuda"))


@pytest.mark.gpu
def test_clip_matches_reference_on_cuda():
    print("clip (fwb200 kernels):", _check_clip("cuda"))


@pytest.mark.gpu
def test_pipeline_encode_image_on_cuda():
    _check_pipeline_encode_image("cuda")


def _full_size(model, dtype=BF16):
    from fwb_synth import materialize, synth_init
    return synth_init(materialize(model, "cuda", dtype), seed=0).eval()


@pytest.mark.gpu
def test_full_size_umt5_kernels_vs_fp32_accumulating_shim():
    """umT5-XXL (24 x [4096, 64 heads, 10240]) at 512 tokens with a 37-token prompt: the fwb200 kernels against the torch shim on
    the same weights (both round where the ABI says; the shim accumulates in fp32 with torch's summation order)."""
    from FantasyWorld.diffsynth_wan21.models.wan_video_text_encoder import WanTextEncoder, init_weights
    from fwb_synth import materialize
    with torch.device("meta"):
        m = WanTextEncoder()
    # the encoder's own initialisation (wan_video_text_encoder.py:192-208: q at std 1/dim, i.e. soft attention), not the per-key
    # synthetic one: T5 has no 1/sqrt(d) in the logits, and unit-variance q AND k give near one-hot softmaxes whose arg-max flips
    # under bf16 rounding — 24 such layers decorrelate any two correct implementations (measured: 0.86)
    torch.manual_seed(0)
    m = materialize(m, "cuda", BF16).eval()
    with torch.no_grad():
        m.apply(init_weights)
        torch.nn.init.normal_(m.token_embedding.weight)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, 256384, (1, 512), generator=g).cuda()
    mask = torch.zeros(1, 512, dtype=torch.long, device="cuda")
    mask[:, :37] = 1
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = m(ids, mask)
    ev0.record()
    out = m(ids, mask)
    ev1.record()
    torch.cuda.synchronize()
    with torch_ops():
        ref = m(ids, mask)
    assert out.shape == (1, 512, 4096) and torch.isfinite(out.float()).all()
    e = rel_err(out[:, :37], ref[:, :37])
    print(f"umT5-XXL forward {ev0.elapsed_time(ev1):.1f} ms; kernels vs shim rel err {e:.3e}")
    assert e < 5e-2
    # block by block on the SAME input (no error propagation): every kernel shape of the encoder against the shim
    x = m.token_embedding(ids).to(BF16)
    worst = 0.0
    for blk in m.blocks:
        y = blk(x, mask)
        with torch_ops():
            yr = blk(x, mask)
        worst = max(worst, rel_err(y[:, :37], yr[:, :37]))
        x = y
    print(f"umT5-XXL worst single-block rel err (same input): {worst:.3e}")
    assert worst < 1e-2


@pytest.mark.gpu
def test_full_size_clip_kernels_vs_fp32_accumulating_shim():
    """ViT-H/14 (31 of 32 blocks, 257 tokens, head_dim 80 on the head_dim-96 attention instance) through encode_image."""
    from FantasyWorld.diffsynth_wan21.models.wan_video_image_encoder import WanImageEncoder
    enc = _full_size(WanImageEncoder(device="meta"))
    g = torch.Generator().manual_seed(4)
    img = (torch.rand(1, 3, 480, 832, generator=g) * 2 - 1).cuda()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = enc.encode_image([img])
    ev0.record()
    out = enc.encode_image([img])
    ev1.record()
    torch.cuda.synchronize()
    with torch_ops():
        ref = enc.encode_image([img])
    assert out.shape == (1, 257, 1280) and out.dtype == BF16 and torch.isfinite(out.float()).all()
    e = rel_err(out, ref)
    print(f"CLIP ViT-H encode_image {ev0.elapsed_time(ev1):.1f} ms; kernels vs shim rel err {e:.3e}")
    assert e < 5e-2
    x, worst = ref, 0.0
    for blk in enc.model.visual.transformer:
        y = blk(x)
        with torch_ops():
            yr = blk(x)
        worst = max(worst, rel_err(y, yr))
        x = y
    print(f"CLIP ViT-H worst single-block rel err (same input): {worst:.3e}")
    assert worst < 1e-2
