"""The Wan2.2 conditioning call `pipe(prompt=..., negative_prompt=..., input_image=..., end_image=..., seed=..., tiled=True, ...,
return_condition=True)` (inference_wan22.py:345-353) and the tiled VAE encode behind it, against a golden written by the UNMODIFIED
reference pipeline on the CPU (tools/make_golden_wan22_cond.py).  CPU: VAE in fp32 (torch), umT5 host logic on the test-only torch shim;
GPU (-m gpu): the same call with the umT5 on the fwb200 kernels."""
import types

import pytest
import torch

from _common import gold, rel_err
from _ops_torch_shim import torch_ops
from test_encoders import FakeTokenizer, _t5

BF16 = torch.bfloat16


def _vae(device, dtype=torch.float32):
    from FantasyWorld.diffsynth_wan22.models.wan_video_vae import WanVideoVAE
    from fwb_synth import synth_init
    wrap = torch.nn.Module()
    wrap.vae = WanVideoVAE(z_dim=16)
    wrap.vae.model.requires_grad_(True)
    synth_init(wrap, seed=0, gen_device="cpu")
    wrap.vae.model.requires_grad_(False)
    return wrap.vae.to(device=device, dtype=dtype).eval()


def _pipe(device):
    from FantasyWorld.diffsynth_wan22.pipelines.wan_video_new import WanVideoPipeline
    g = gold("wan22_condition.pt")
    pipe = WanVideoPipeline(device=device, torch_dtype=torch.float32)
    pipe.text_encoder = _t5(g["t5_cfg"], device)
    pipe.prompter.fetch_models(pipe.text_encoder)
    pipe.prompter.tokenizer = FakeTokenizer(24, g["t5_cfg"]["vocab"])
    pipe.vae = _vae(device)
    pipe.height_division_factor = pipe.width_division_factor = pipe.vae.upsampling_factor * 2
    pipe.dit = types.SimpleNamespace(require_vae_embedding=True)
    return pipe, g


def _check_call(device):
    from PIL import Image
    pipe, g = _pipe(device)
    a, b = Image.fromarray(g["pil_a"].numpy()), Image.fromarray(g["pil_b"].numpy())
    for tag, end in (("first", None), ("first_last", b)):
        ref = g[tag]
        shared, posi, nega = pipe(input_image=a, end_image=end, **g["call"])
        assert (shared["height"], shared["width"], shared["num_frames"]) == (ref["height"], ref["width"], ref["num_frames"]) == (32, 48, 9)
        assert torch.equal(shared["noise"].cpu(), ref["noise"]) and shared["latents"] is shared["noise"]          # CPU generator, seed 3
        assert torch.equal(pipe.scheduler.timesteps.cpu(), ref["timesteps"])
        y = shared["y"].cpu()
        assert y.shape == ref["y"].shape == (1, 20, 3, 4, 6) and torch.equal(y[:, :4], ref["y"][:, :4])          # known-frame mask: exact
        tol = 2e-5 if device == "cpu" else 2e-3                # fp32 VAE; cuDNN may pick TF32 convolutions, as for the reference on a GPU
        assert rel_err(y[:, 4:], ref["y"][:, 4:]) < tol, rel_err(y[:, 4:], ref["y"][:, 4:])
        for ours, theirs in ((posi["context"], ref["context_pos"]), (nega["context"], ref["context_neg"])):
            ours = ours.cpu()
            assert ours.shape == theirs.shape and torch.equal((ours == 0).all(-1), (theirs == 0).all(-1))      # same padding rows zeroed
            assert rel_err(ours, theirs) < 0.12                 # reduced umT5 in bf16 vs the fp32 golden (budget: tests/test_encoders.py)
    with pytest.raises(NotImplementedError, match="conditioning only"):
        pipe(prompt="x", return_condition=False)
    with pytest.raises(NotImplementedError, match="outside the FantasyWorld path"):
        pipe(prompt="x", return_condition=True, control_video=[a])


def test_wan22_condition_call_matches_reference():
    with torch_ops():
        _check_call("cpu")


@pytest.mark.gpu
def test_wan22_condition_call_on_cuda():
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # fp32 VAE convolutions in fp32, so that the comparison with the CPU golden is tight
    try:
        _check_call("cuda")
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def test_tiled_vae_encode_matches_reference():
    g = gold("wan22_condition.pt")
    vae = _vae("cpu")
    with torch.no_grad():
        tiled = vae.encode([g["clip"]], device="cpu", tiled=True, tile_size=(3, 4), tile_stride=(2, 2))
        single = vae.encode([g["clip"]], device="cpu", tiled=False)
    assert tiled.shape == g["enc_tiled"].shape == (1, 16, 2, 4, 6)
    assert rel_err(tiled, g["enc_tiled"]) < 2e-5 and rel_err(single, g["enc_single"]) < 2e-5
    assert rel_err(g["enc_tiled"], g["enc_single"]) > 1e-2          # tiling really changes the result: the comparison is sharp


def test_wan22_pipeline_helpers_match_reference():
    from PIL import Image
    from FantasyWorld.diffsynth_wan22.pipelines.wan_video_new import ModelConfig, WanVideoPipeline
    g = gold("wan22_condition.pt")
    pipe = WanVideoPipeline(device="cpu", torch_dtype=BF16)
    a = Image.fromarray(g["pil_a"].numpy())
    assert torch.equal(pipe.preprocess_image(a.resize((48, 32))), g["preprocess_bf16"])      # scaling in bf16, like the reference
    assert torch.equal(pipe.generate_noise((1, 16, 3, 4, 6), seed=5), g["noise_bf16"])      # fp32 CPU draw, then bf16
    assert pipe.check_resize_height_width(480, 832, 81) == (480, 832, 81)
    assert pipe.check_resize_height_width(481, 830, 82) == (496, 832, 85) and pipe.check_resize_height_width(30, 48) == (32, 48)
    with pytest.raises(RuntimeError, match="text encoder"):
        pipe(prompt="x", return_condition=True)
    assert ModelConfig(model_id="m", origin_file_pattern="*.pth", local_model_path="/nonexistent").local_files() == []


def test_wan22_from_pretrained_resolves_local_side_checkpoints(tmp_path):
    """fusion/model_wan22.py with load_vae / load_text_encoder: the expert's pipeline is built from the DiT pattern plus
    `models_t5_umt5-xxl-enc-bf16.pth` and `Wan2.1_VAE.pth` under the same local directory (ref: model_wan22.py:144-163)."""
    from FantasyWorld.diffsynth_wan21.models.wan_video_vae import WanVideoVAE
    from FantasyWorld.diffsynth_wan22.pipelines.wan_video_new import ModelConfig, WanVideoPipeline
    g = gold("wan22_condition.pt")
    t5 = _t5(g["t5_cfg"])
    torch.save(t5.state_dict(), tmp_path / "models_t5_umt5-xxl-enc-bf16.pth")
    torch.save({k[len("model."):]: v for k, v in WanVideoVAE(z_dim=16).state_dict().items()}, tmp_path / "Wan2.1_VAE.pth")
    (tmp_path / "google" / "umt5-xxl").mkdir(parents=True)
    cfgs = [ModelConfig(model_id="PAI/x", origin_file_pattern=p, local_model_path=str(tmp_path))
            for p in ("high_noise_model/diffusion_pytorch_model*.safetensors", "models_t5_umt5-xxl-enc-bf16.pth", "Wan2.1_VAE.pth")]
    tiny_dit = dict(dim=64, in_dim=36, ffn_dim=128, out_dim=16, text_dim=64, freq_dim=32, eps=1e-6, patch_size=(1, 2, 2), num_heads=2,
                    num_layers=1, has_image_input=False)
    pipe = WanVideoPipeline.from_pretrained(torch_dtype=BF16, device="cpu", model_configs=cfgs, tokenizer_config=None, dit_config=tiny_dit,
                                            side_configs={"wan_video_text_encoder": g["t5_cfg"]})
    assert pipe.text_encoder is not None and pipe.vae is not None and pipe.image_encoder is None and pipe.dit is not None
    assert pipe.height_division_factor == 16 and pipe.prompter.tokenizer is None
    a, b = pipe.text_encoder.state_dict(), t5.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
