"""Synthetic FantasyWorldFusionModel for tests and benchmarks.  The generators themselves live in the torch-only module
`fwb_synth` (no native library); this module adds the model builder, which needs the kernel-backed mirror."""
from __future__ import annotations

import torch
import torch.nn as nn

from fwb_synth import (CAMERA_CFG, VGGT_CFG, WAN21_I2V_14B, materialize, synth_init, synth_inputs,  # noqa: F401
                       synth_tensor)


def build_fusion_model(num_dit_layers: int = 40, start_index: int = 16, device="cuda", seed: int = 0, heads: bool = True,
                       gen_device=None):
    """Random-init FantasyWorldFusionModel (Wan2.1-I2V-14B widths; depth reducible for tests) with synthetic weights,
    built on the meta device so the 18.5 B parameters are never materialised on the host."""
    from FantasyWorld.diffsynth_wan21.models.wan_video_dit import precompute_freqs_cis_3d
    from FantasyWorld.fusion.model_wan21 import FantasyWorldFusionModel
    from FantasyWorld.wan.modules.model import rope_params

    n_irg = num_dit_layers - start_index
    cfg = dict(WAN21_I2V_14B, num_layers=num_dit_layers)
    vcfg = dict(VGGT_CFG)
    if not heads:
        vcfg.update(enable_camera=False, enable_depth=False, enable_point=False)
    with torch.device("meta"):
        model = FantasyWorldFusionModel(start_index=start_index, use_gradient_checkpointing=False,
                                        cross_attention_list=list(range(n_irg)), dit_path=None, vggt_cfg=vcfg,
                                        camera_control=True, camera_cfg=dict(CAMERA_CFG), dit_config=cfg, device=str(device))
    agg = model.vggt.aggregator
    agg.frame_blocks = nn.ModuleList(list(agg.frame_blocks)[:n_irg])
    agg.global_blocks = nn.ModuleList(list(agg.global_blocks)[:n_irg])
    materialize(model, device, torch.bfloat16)
    # plain (non-buffer) tables created under the meta device: rebuild them for real
    model.pipe.dit.freqs = precompute_freqs_cis_3d(cfg["dim"] // cfg["num_heads"])
    model.freqs_bicross = precompute_freqs_cis_3d(model.bicross_dim // model.bicross_num_heads)
    d = 1024 // 16
    agg.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1)
    from FantasyWorld.fusion.model_wan21 import LATENT_MEAN, LATENT_STD
    model.mean, model.std = torch.tensor(LATENT_MEAN), torch.tensor(LATENT_STD)
    model.scale = [model.mean, 1.0 / model.std]
    model.pipe.device = str(device)
    model.pipe.torch_dtype = torch.bfloat16
    synth_init(model, seed, gen_device=gen_device)
    return model.eval()


def build_fusion_model_wan22(num_dit_layers: int = 40, start_index: int = 16, device="cuda", seed: int = 0, heads: bool = False,
                             gen_device=None):
    """Random-init Wan2.2-Fun-A14B-Control-Camera fusion model (BASELINE config 4 family; one of the two experts), same
    construction as build_fusion_model: meta device, materialise in bf16, per-key synthetic weights."""
    from FantasyWorld.diffsynth_wan21.models.wan_video_dit import precompute_freqs_cis_3d
    from FantasyWorld.diffsynth_wan22.models.wan_video_dit import WAN22_FUN_A14B_CONTROL_CAMERA
    from FantasyWorld.fusion.model_wan22 import FantasyWorldFusionModel as Fusion22
    from FantasyWorld.wan.modules.model import rope_params

    n_irg = num_dit_layers - start_index
    vcfg = dict(VGGT_CFG)
    if not heads:
        vcfg.update(enable_camera=False, enable_depth=False, enable_point=False)
    with torch.device("meta"):
        model = Fusion22(start_index=start_index, use_gradient_checkpointing=False, cross_attention_list=list(range(n_irg)), dit_path=None,
                         lora_path=None, vggt_cfg=vcfg, camera_control=True, camera_cfg=dict(use_info="plucker"),
                         dit_config=dict(WAN22_FUN_A14B_CONTROL_CAMERA, num_layers=num_dit_layers))
    agg = model.vggt.aggregator
    agg.frame_blocks = nn.ModuleList(list(agg.frame_blocks)[:n_irg])
    agg.global_blocks = nn.ModuleList(list(agg.global_blocks)[:n_irg])
    materialize(model, device, torch.bfloat16)
    model.pipe.dit.freqs = precompute_freqs_cis_3d(128)
    model.freqs_bicross = precompute_freqs_cis_3d(model.bicross_dim // model.bicross_num_heads)
    d = 1024 // 16
    agg.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1)
    model.pipe.device = str(device)
    model.pipe.torch_dtype = torch.bfloat16
    synth_init(model, seed, gen_device=gen_device)
    return model.eval()


def build_vggt(device="cuda", seed: int = 0, heads: bool = True, gen_device=None):
    """Random-init stand-alone geometry branch (BASELINE config 5): VGGT with its 24 frame + 24 global blocks and the heads."""
    from FantasyWorld.vggt.models.vggt import VGGT
    from FantasyWorld.wan.modules.model import rope_params

    vcfg = dict(VGGT_CFG)
    if not heads:
        vcfg.update(enable_camera=False, enable_depth=False, enable_point=False)
    with torch.device("meta"):
        wrap = nn.Module()
        wrap.vggt = VGGT(**vcfg)          # keys `vggt.*` as inside the fusion model, so the per-key weights are the same tensors
    materialize(wrap, device, torch.bfloat16)
    agg = wrap.vggt.aggregator
    d = 1024 // 16
    agg.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)), rope_params(1024, 2 * (d // 6))], dim=1)
    synth_init(wrap, seed, gen_device=gen_device)
    return wrap.vggt.eval()
