"""Golden vector for the Wan2.2-Fun-A14B-Control-Camera fusion model: the UNMODIFIED reference
FantasyWorld.fusion.model_wan22.FantasyWorldFusionModel.joint_forward, reduced depth (1 PCB + 1 IRG, 14B widths), CPU fp32,
f,h,w = 2,4,4, with the control adapter and without heads.   python tools/make_golden_wan22.py
"""
import copy
import json
import sys
import types
from pathlib import Path

import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))
from ref_shim import VGGT_CFG, install_stubs, randomize_zero_init

install_stubs()
import importlib

for name in ("modelscope", "imageio"):
    pass
try:
    fusion = importlib.import_module("FantasyWorld.fusion.model_wan22")
except Exception as e:  # the wan22 pipeline pulls more optional deps; stub what is missing and retry
    print("first import failed:", repr(e))
    raise
dit22 = importlib.import_module("FantasyWorld.diffsynth_wan22.models.wan_video_dit")
dit22.FLASH_ATTN_2_AVAILABLE = False
dit22.FLASH_ATTN_3_AVAILABLE = False
dit22.SAGE_ATTN_AVAILABLE = False
blockmod = importlib.import_module("FantasyWorld.fusion.layer.block")
dit21 = importlib.import_module("FantasyWorld.diffsynth_wan21.models.wan_video_dit")
dit21.FLASH_ATTN_2_AVAILABLE = False
dit21.FLASH_ATTN_3_AVAILABLE = False
dit21.SAGE_ATTN_AVAILABLE = False
vggt_mod = importlib.import_module("FantasyWorld.vggt.models.vggt")
from fwb_synth import synth_init, synth_inputs

torch.manual_seed(0)
torch.set_grad_enabled(False)
model = fusion.FantasyWorldFusionModel.__new__(fusion.FantasyWorldFusionModel)
nn.Module.__init__(model)


class _Pipe(nn.Module):
    pass


pipe = _Pipe()
pipe.dit = dit22.WanModel(dim=5120, in_dim=36, ffn_dim=13824, out_dim=16, text_dim=4096, freq_dim=256, eps=1e-6, patch_size=(1, 2, 2),
                          num_heads=40, num_layers=2, has_image_input=False, add_control_adapter=True, in_dim_control_adapter=24,
                          require_clip_embedding=False)
model.pipe = pipe
cfg = dict(VGGT_CFG, enable_camera=False, enable_depth=False, enable_point=False)
model.vggt = vggt_mod.VGGT(**cfg)
model.vggt.aggregator.frame_blocks = nn.ModuleList(list(model.vggt.aggregator.frame_blocks)[:1])
model.vggt.aggregator.global_blocks = nn.ModuleList(list(model.vggt.aggregator.global_blocks)[:1])
model.start_index, model.cross_attention_list = 1, [0]
model.freqs_bicross = dit22.precompute_freqs_cis_3d(96)
src_dit, src_agg = pipe.dit.blocks[1], model.vggt.aggregator.global_blocks[0]
d, a = copy.deepcopy(src_dit), copy.deepcopy(src_agg)
pipe.dit.blocks[1] = nn.Identity()
model.vggt.aggregator.global_blocks[0] = nn.Identity()
model.IRGBlock = nn.ModuleList([blockmod.IRGBlock(x_dit_block=d, x_agg_block=a, m1_dim=5120, m2_dim=1024, hidden_size=1152, num_heads=12, drop_path=None)])
synth_init(model, seed=0, gen_device="cpu")
model.eval()
schema = {k: list(v.shape) for k, v in model.state_dict().items()}
GOLD = ROOT / "tests" / "golden"
(GOLD / "schema_wan22_reduced.json").write_text(json.dumps(schema, indent=0))

F, H, W = 2, 4, 4
inp = synth_inputs(F, H, W, device="cpu", seed=1024, text_len=64, dtype=torch.float32)
g = torch.Generator().manual_seed(99)
control = torch.randn(1, 24, F, 16 * H, 16 * W, generator=g)
out, _ = model.joint_forward(inp["latents"], timestep=torch.tensor([996.0]), context=inp["context_pos"], y=inp["y"],
                             use_gradient_checkpointing=False, control_camera_latents_input=control)
print(out.shape, out.abs().mean().item())
torch.save({"out": out.clone(), "grid": (F, H, W), "text_len": 64, "timestep": 996.0, "control_seed": 99}, GOLD / "joint_forward_wan22.pt")
