"""B200-native mirror of FantasyWorld/vggt/models/vggt.py (reference): the geometry branch wrapper — 5120->1024 token
projection, fp32 timestep modulation e0, aggregator, and the camera / depth / point heads.  Same state_dict keys.
(`PyTorchModelHubMixin` of the reference only adds from_pretrained/push_to_hub; the track head is disabled on this
path — inference_wan21.py:193 — and out of scope)
"""
from __future__ import annotations

import torch
import torch.nn as nn

from fwb200 import engine as E
from fwb200 import ops

from ...wan.modules.model import sinusoidal_embedding_1d
from ..heads.camera_head import CameraHead
from ..heads.dpt_head import DPTHead_3D_Causal
from ..models.aggregator import Aggregator


class VGGT(nn.Module):
    def __init__(self, img_size=518, patch_size=16, embed_dim=1024, number_frame=81, freq_dim=256, enable_camera=True,
                 enable_depth=True, enable_point=True, enable_track=True, load_path=None, DPT_patch_size=16):
        super().__init__()
        self.spatial_frame = (number_frame - 1) // 4 + 1
        self.freq_dim, self.embed_dim = freq_dim, embed_dim
        self.projection_head = nn.Conv3d(5120, 1024, kernel_size=(1, 1, 1), stride=(1, 1, 1))
        self.aggregator = Aggregator(img_size=img_size, patch_size=patch_size, embed_dim=embed_dim, spatial_time=self.spatial_frame)
        self.camera_head = CameraHead(dim_in=2 * embed_dim) if enable_camera else None
        self.depth_head = DPTHead_3D_Causal(dim_in=2 * embed_dim, output_dim=2, activation="exp", conf_activation="expp1",
                                            patch_size=DPT_patch_size) if enable_depth else None
        self.point_head = DPTHead_3D_Causal(dim_in=2 * embed_dim, output_dim=4, activation="inv_log", conf_activation="expp1",
                                            patch_size=DPT_patch_size) if enable_point else None
        if enable_track:
            raise NotImplementedError("track head is out of scope (enable_track=False on the FantasyWorld path)")
        self.track_head = None
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, embed_dim), nn.SiLU(), nn.Linear(embed_dim, embed_dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(embed_dim, embed_dim * 6))
        if load_path is not None:
            self.load_state_dict(torch.load(load_path)['model'], strict=True)

    # ---- pieces ------------------------------------------------------------------------------------------------------
    def project_tokens(self, tokens: torch.Tensor) -> torch.Tensor:
        """1x1x1 Conv3d(5120->1024) as a GEMM over tokens [..., 5120] -> [..., 1024].  ref: vggt.py:32, 123."""
        return E.lin(E.as_bf16(tokens).reshape(-1, tokens.shape[-1]), self.projection_head,
                     round_flags=ops.ROUND_AFTER_BIAS).view(*tokens.shape[:-1], -1)

    def time_modulation(self, t: torch.Tensor) -> torch.Tensor:
        """e0 [B, 6, 1024] in fp32 (the reference forces fp32 here).  6 MFLOP: plain torch.  ref: vggt.py:126-130."""
        with torch.autocast(device_type=t.device.type, enabled=False):
            s = sinusoidal_embedding_1d(self.freq_dim, t).float()
            f = torch.nn.functional
            l0, l2, lp = self.time_embedding[0], self.time_embedding[2], self.time_projection[1]
            e = f.linear(f.silu(f.linear(s, l0.weight.float(), l0.bias.float())), l2.weight.float(), l2.bias.float())
            e0 = f.linear(f.silu(e), lp.weight.float(), lp.bias.float()).unflatten(1, (6, self.embed_dim))
        return e0

    def _process_wan_input(self, patch_token: torch.Tensor, query_points: torch.Tensor = None,
                           camera_token: torch.Tensor = None, t=None):
        """patch_token [B, 5120, T, h, w] -> ([B, T, h, w, 1024], camera_token, e0).  ref: vggt.py:118-131."""
        tok = self.project_tokens(patch_token.permute(0, 2, 3, 4, 1))
        return tok, camera_token, self.time_modulation(t)

    def _head_predction(self, patch_token, patch_start_idx, aggregated_tokens_list):
        """ref: vggt.py:134-154."""
        out = {}
        with torch.autocast(device_type="cuda", enabled=True, dtype=torch.bfloat16):
            if self.camera_head is not None:
                out["pose_enc"] = self.camera_head(aggregated_tokens_list)[-1]
            if self.depth_head is not None:
                out["depth"], out["depth_conf"] = self.depth_head(aggregated_tokens_list, images=patch_token, patch_start_idx=patch_start_idx)
            if self.point_head is not None:
                out["world_points"], out["world_points_conf"] = self.point_head(aggregated_tokens_list, images=patch_token, patch_start_idx=patch_start_idx)
        return out

    def forward(self, patch_token: torch.Tensor, query_points: torch.Tensor = None, camera_token: torch.Tensor = None, t=None):
        """Stand-alone geometry branch (BASELINE config 5).  ref: vggt.py:45-117."""
        tok, camera_token, e0 = self._process_wan_input(patch_token, query_points, camera_token, t)
        tokens_list, patch_start_idx = self.aggregator(tok, camera_token, e0)
        return self._head_predction(tok, patch_start_idx, tokens_list)
