"""Golden vector for the camera pose encoder: the UNMODIFIED reference CameraPoseEncoder on CPU fp32 with the
per-key synthetic weights, input = seeded Plücker embedding [1, 5, 64, 64, 6] (5 video frames -> 2 latent frames, 4x4 tokens).
    python tools/make_golden_pose.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

from ref_shim import install_stubs
install_stubs()
import importlib
from fwb_synth import synth_init

mod = importlib.import_module("FantasyWorld.diffsynth_wan21.models.pose_adaptor_ac3d")
torch.manual_seed(0)
enc = mod.CameraPoseEncoder(context_dim=2048, in_channels=6, downscale_coef=8, pose_inject_method="adaln")
wrap = torch.nn.Module()
wrap.camera_condition = torch.nn.Module()
wrap.camera_condition.pose_encoder = enc
synth_init(wrap, seed=0, gen_device="cpu")
g = torch.Generator().manual_seed(77)
x = torch.randn(1, 5, 64, 64, 6, generator=g)
with torch.no_grad():
    y = enc.eval()(x)
print(y.shape, y.abs().mean().item())
torch.save({"out": y.clone(), "seed": 77, "schema": {k: list(v.shape) for k, v in wrap.state_dict().items()}},
           ROOT / "tests" / "golden" / "pose_encoder.pt")
