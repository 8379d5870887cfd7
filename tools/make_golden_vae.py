"""Golden for the Wan VAE mirror (SURVEY §8f N1 tiled decode, N3 encode): the UNMODIFIED reference WanVideoVAE
(FantasyWorld/diffsynth_wan21/models/wan_video_vae.py) on CPU fp32 with the per-key synthetic weights.

    python tools/make_golden_vae.py      # build container -> tests/golden/vae.pt (~0.6 MB: schema + small outputs)

Cases: untiled decode of a [1,16,3,6,8] latent (-> 9 frames of 48x64), tiled decode with overlapping tiles (tile 4x6, stride 2x3:
interior + border tiles, ramps on all four sides), encode of a 9-frame 32x48 clip.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))


def inputs():
    g = torch.Generator().manual_seed(11)
    return dict(z=torch.randn(1, 16, 3, 6, 8, generator=g), video=torch.randn(3, 9, 32, 48, generator=g).clamp(-1, 1))


def main():
    import importlib
    from ref_shim import import_reference
    from fwb_synth import synth_init
    import_reference()
    ref = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_vae")
    torch.manual_seed(0)
    wrap = torch.nn.Module()
    wrap.vae = ref.WanVideoVAE(z_dim=16)
    wrap.vae.model.requires_grad_(True)           # synth_init walks named_parameters
    synth_init(wrap, seed=0, gen_device="cpu")
    vae = wrap.vae.eval()
    inp = inputs()
    with torch.no_grad():
        single = vae.decode(inp["z"], device="cpu", tiled=False)
        tiled = vae.decode(inp["z"], device="cpu", tiled=True, tile_size=(4, 6), tile_stride=(2, 3))
        enc = vae.encode([inp["video"]], device="cpu", tiled=False)
    out = {"schema": {k: list(v.shape) for k, v in wrap.state_dict().items()}, "single": single, "tiled": tiled, "enc": enc,
           "tile_size": (4, 6), "tile_stride": (2, 3)}
    path = ROOT / "tests" / "golden" / "vae.pt"
    torch.save(out, path)
    print("wrote", path, path.stat().st_size, "bytes", single.shape, tiled.shape, enc.shape, len(out["schema"]), "keys")


if __name__ == "__main__":
    main()
