"""One whole sample through the public API on ONE B200 — the `sec/video` half of BASELINE.json's metric (81 frames, 480x832, 14B):

    prompt (x2: positive / negative)  -> WanVideoPipeline.encode_prompt      (umT5-XXL mirror; stand-in tokenizer, no files offline)
    first frame                       -> WanVideoPipeline.encode_image       (CLIP ViT-H mirror + Wan VAE encoder)
    camera rays [1, 81, 480, 832, 6]  -> FantasyWorldFusionModel.generate_video: pose encoder, 50 CFG steps (2 forwards each),
                                         geometry heads (depth / points / camera / confidences at 81 x 480 x 832) on the last step
    final latents                     -> WanVideoVAE.decode(tiled=True, tile (30, 52), stride (15, 26))   (inference_wan21.py:324-330)

Random-init weights of the real architectures (no checkpoints here), synthetic image / rays / prompts.  Every stage is timed with a
device synchronize on both sides and logged as soon as it finishes (a later failure keeps the earlier numbers); cheap stages run first.

    python tools/gpu_video_e2e.py [--steps 50] [--out gpurun_out/r02_video_e2e.json]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

import torch  # noqa: E402


class WordHashTokenizer:
    """Stand-in for the umT5 tokenizer (its files are not available offline): whitespace words -> ids, right-padded to 512 with a
    mask — the interface WanPrompter.encode_prompt uses (prompters/wan_prompter.py)."""

    def __init__(self, seq_len=512, vocab=256384):
        self.seq_len, self.vocab = seq_len, vocab

    def __call__(self, sequence, return_mask=False, add_special_tokens=True):
        if isinstance(sequence, str):
            sequence = [sequence]
        ids = torch.zeros(len(sequence), self.seq_len, dtype=torch.long)
        mask = torch.zeros_like(ids)
        for i, s in enumerate(sequence):
            toks = [2 + hash_word(w) % (self.vocab - 2) for w in s.split()][: self.seq_len - 1] + [1]
            ids[i, :len(toks)] = torch.tensor(toks)
            mask[i, :len(toks)] = 1
        return (ids, mask) if return_mask else ids


def hash_word(w: str) -> int:
    h = 2166136261
    for ch in w.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


class Log:
    def __init__(self, path):
        self.path, self.rec = Path(path), {"stages": {}, "errors": {}}
        self.path.parent.mkdir(parents=True, exist_ok=True)

    def stage(self, name, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            out = fn()
            torch.cuda.synchronize()
            self.rec["stages"][name] = round(time.perf_counter() - t0, 4)
            print(f"[e2e] {name}: {self.rec['stages'][name]:.3f} s", flush=True)
        except Exception as e:  # keep going: later stages have synthetic stand-ins for a failed stage's outputs
            torch.cuda.synchronize()
            out = None
            self.rec["errors"][name] = f"{type(e).__name__}: {e}"
            print(f"[e2e] {name} FAILED: {type(e).__name__}: {e}", flush=True)
            traceback.print_exc()
        self.flush()
        return out

    def flush(self):
        self.path.write_text(json.dumps(self.rec, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--frames", type=int, default=81)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=832)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "r02_video_e2e.json"))
    a = ap.parse_args()
    import fwb200
    from fwb200.synth import build_fusion_model
    from PIL import Image
    fwb200.require_device()
    dev = "cuda"
    log = Log(a.out)
    log.rec.update(config=dict(steps=a.steps, frames=a.frames, height=a.height, width=a.width, model="Wan2.1-I2V-14B shape, 16 PCB + 24 IRG, "
                               "umT5-XXL, CLIP ViT-H/14, Wan VAE; random init", device=torch.cuda.get_device_name(0)))
    f_lat = (a.frames - 1) // 4 + 1

    model = log.stage("build_fusion_model", lambda: build_fusion_model(40, 16, device=dev, seed=0, heads=True))
    pipe = model.pipe
    log.stage("build_text_encoder", lambda: pipe.enable_text_encoder(device=dev))
    log.stage("build_image_encoder", lambda: pipe.enable_image_encoder(device=dev))
    log.stage("build_vae", lambda: pipe.enable_vae(device=dev))
    pipe.prompter.tokenizer = WordHashTokenizer()
    log.rec["memory_gb_after_build"] = round(torch.cuda.memory_allocated() / 2 ** 30, 2)

    g = torch.Generator().manual_seed(7)
    image = Image.fromarray(torch.randint(0, 256, (a.height, a.width, 3), generator=g, dtype=torch.uint8).numpy())
    rays = torch.randn(1, a.frames, a.height, a.width, 6, generator=g).to(device=dev, dtype=torch.bfloat16)
    pos_text = "a slow dolly shot through a sunlit museum hall, marble statues on both sides, dust in the light beams"
    neg_text = "blurry, low quality, distorted, static frame, watermark, text, jpeg artifacts"

    # ---- conditioning (once per sample) --------------------------------------------------------------------------------------
    for warm in (True, False):          # the first call of each encoder builds its cached fused / padded weights
        tag = "_first_call" if warm else ""
        cpos = log.stage("encode_prompt_positive" + tag, lambda: pipe.encode_prompt(pos_text, positive=True)["context"])
        cneg = log.stage("encode_prompt_negative" + tag, lambda: pipe.encode_prompt(neg_text, positive=False)["context"])
    img = log.stage("encode_image", lambda: pipe.encode_image(image, None, a.frames, a.height, a.width))
    if cpos is None or cneg is None:
        cpos = torch.randn(1, 512, 4096, device=dev, dtype=torch.bfloat16)
        cneg = torch.randn(1, 512, 4096, device=dev, dtype=torch.bfloat16)
    if img is None:
        img = {"clip_feature": torch.randn(1, 257, 1280, device=dev, dtype=torch.bfloat16),
               "y": torch.randn(1, 20, f_lat, a.height // 8, a.width // 8, device=dev, dtype=torch.bfloat16)}
        log.rec["errors"]["encode_image_fallback"] = "synthetic clip_feature / y used for the sampler"
    log.rec["shapes"] = {"context": list(cpos.shape), "clip_feature": list(img["clip_feature"].shape), "y": list(img["y"].shape)}

    def sample(steps):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model.generate_video(context_pos=cpos, context_neg=cneg, clip_feature=img["clip_feature"], y=img["y"],
                                        height=a.height, width=a.width, num_frames=a.frames, num_inference_steps=steps, cfg_scale=5.0,
                                        seed=0, device=dev, plucker_embedding=rays)

    def decode(lat):
        with torch.no_grad():
            return pipe.vae.decode(lat, device=dev, tiled=True, tile_size=(30, 52), tile_stride=(15, 26))

    # ---- a 2-step dry run of the sampler + heads before the long loop --------------------------------------------------------
    dry = log.stage("dry_run_2_steps_with_heads", lambda: sample(2))
    if dry is not None:
        lat2, pred2 = dry
        log.rec["prediction_keys"] = {k: list(v.shape) for k, v in pred2.items() if torch.is_tensor(v)} if pred2 else None
        del lat2, pred2, dry          # (the tiled decode at this size is exercised by tests/test_gpu_parity.py)
        torch.cuda.empty_cache()

    # ---- the sample -----------------------------------------------------------------------------------------------------------
    fwb200.reset_launch_count()
    res = log.stage(f"generate_video_{a.steps}_steps_with_heads", lambda: sample(a.steps))
    log.rec["fwb200_launches_in_sampler"] = fwb200.launch_count()
    if res is not None:
        latents, pred = res
        log.rec["latents_finite"] = bool(torch.isfinite(latents.float()).all())
        video = log.stage("tiled_vae_decode", lambda: decode(latents))
        if video is not None:
            log.rec["video_finite"] = bool(torch.isfinite(video.float()).all())
    st = log.rec["stages"]
    parts = ["encode_prompt_positive", "encode_prompt_negative", "encode_image", f"generate_video_{a.steps}_steps_with_heads",
             "tiled_vae_decode"]
    if all(p in st for p in parts):
        log.rec["sec_per_video"] = round(sum(st[p] for p in parts), 3)
        log.rec["sec_per_video_parts"] = {p: st[p] for p in parts}
        log.rec["denoise_steps_per_sec_in_sampler"] = round(a.steps / st[parts[3]], 4)
    log.rec["peak_memory_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    log.flush()
    print(json.dumps(log.rec))


if __name__ == "__main__":
    main()
