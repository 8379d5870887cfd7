"""Camera (extrinsics, intrinsics) <-> 9-D pose encoding "absT_quaR_FoV" = translation (3) | quaternion xyzw (4) | vertical and
horizontal field of view (2).  Mirror of FantasyWorld/vggt/utils/pose_enc.py (same names / signatures); used by the reference
CLIs before and after the sampler (inference_wan21.py:18, inference_wan22.py:20), host-side, once per sample."""
from __future__ import annotations

import torch

from .rotation import mat_to_quat, quat_to_mat


def extri_intri_to_pose_encoding(extrinsics, intrinsics, image_size_hw=None, pose_encoding_type="absT_quaR_FoV"):
    """extrinsics [B, S, 3, 4] (camera-from-world, OpenCV), intrinsics [B, S, 3, 3] (pixels) -> [B, S, 9] fp32."""
    if pose_encoding_type != "absT_quaR_FoV":
        raise NotImplementedError
    height, width = image_size_hw
    rot, trans = extrinsics[..., :3, :3], extrinsics[..., :3, 3]
    fov_v = 2 * torch.atan((height / 2) / intrinsics[..., 1, 1])
    fov_h = 2 * torch.atan((width / 2) / intrinsics[..., 0, 0])
    return torch.cat([trans, mat_to_quat(rot), fov_v[..., None], fov_h[..., None]], dim=-1).float()


def pose_encoding_to_extri_intri(pose_encoding, image_size_hw=None, pose_encoding_type="absT_quaR_FoV", build_intrinsics=True):
    """[B, S, 9] -> (extrinsics [B, S, 3, 4], intrinsics [B, S, 3, 3] with the principal point at the image centre, or None)."""
    if pose_encoding_type != "absT_quaR_FoV":
        raise NotImplementedError
    extrinsics = torch.cat([quat_to_mat(pose_encoding[..., 3:7]), pose_encoding[..., :3, None]], dim=-1)
    intrinsics = None
    if build_intrinsics:
        height, width = image_size_hw
        intrinsics = torch.zeros(pose_encoding.shape[:2] + (3, 3), device=pose_encoding.device, dtype=pose_encoding.dtype)
        intrinsics[..., 0, 0] = (width / 2.0) / torch.tan(pose_encoding[..., 8] / 2.0)
        intrinsics[..., 1, 1] = (height / 2.0) / torch.tan(pose_encoding[..., 7] / 2.0)
        intrinsics[..., 0, 2] = width / 2
        intrinsics[..., 1, 2] = height / 2
        intrinsics[..., 2, 2] = 1.0
    return extrinsics, intrinsics
