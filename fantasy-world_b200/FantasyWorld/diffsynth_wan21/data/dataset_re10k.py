"""Mirror of FantasyWorld/diffsynth_wan21/data/dataset_re10k.py: camera trajectories -> per-pixel Pluecker ray embeddings, the input of
the camera-control branch (`plucker_embedding`, inference_wan21.py:284-287 -> fusion/model_wan21.py:267-275 -> CameraPoseEncoder).

Host-side pre-processing, once per sample (numpy float64 camera algebra, float32 torch rays), restated with the reference's order of
operations so that the embedding is bit-identical: `Camera` (:45-56), `create_camera_params_from_batch` (:57-66), `ray_condition`
(:77-119), `RealEstate10KPoseProcessor` (:122-304) — the frame sampling / intrinsics / relative-pose code that the reference repeats in
`get_plucker_embedding` and `get_plucker_embedding_direct_from_cam_params` lives once in `_embed`.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn as nn

from FantasyWorld.vggt.utils.pose_enc import pose_encoding_to_extri_intri


class RandomHorizontalFlipWithPose(nn.Module):
    """All-or-nothing horizontal flip of a clip (training-time augmentation; `use_flip=False` at inference)."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def get_flip_flag(self, n_image):
        keep = torch.rand(1).item() < self.p
        return torch.zeros(n_image, dtype=torch.bool) if keep else torch.ones(n_image, dtype=torch.bool)

    def forward(self, image, flip_flag=None):
        if flip_flag is None:
            flip_flag = self.get_flip_flag(image.shape[0])
        assert image.shape[0] == flip_flag.shape[0]
        return torch.stack([img.flip(-1) if f else img for f, img in zip(flip_flag, image)], dim=0)


class Camera:
    """One RealEstate10K pose line: [id, fx, fy, cx, cy, _, _, w2c (3x4 row-major)] with normalised intrinsics."""

    def __init__(self, entry):
        self.fx, self.fy, self.cx, self.cy = entry[1:5]
        w2c = np.eye(4)
        w2c[:3, :] = np.array(entry[7:]).reshape(3, 4)
        self.w2c_mat = w2c
        self.c2w_mat = np.linalg.inv(w2c)


def create_camera_params_from_batch(extrinsics_np, intrinsics_np):
    """[n, 3, 4] world-to-camera + [n, 3, 3] pixel intrinsics -> Camera list (intrinsics are taken as they are)."""
    cams = []
    for i, (ext, k) in enumerate(zip(extrinsics_np, intrinsics_np)):
        cams.append(Camera([i, k[0, 0], k[1, 1], k[0, 2], k[1, 2], 0, 0] + ext.flatten().tolist()))
    return cams


def custom_meshgrid(*args):
    return torch.meshgrid(*args, indexing="ij")


def _legacy_cross(a, b):
    """`torch.cross(a, b)` as the reference calls it — without `dim`, i.e. along the FIRST dimension of size 3 (the last one for every
    clip that does not have exactly 3 frames or 3 pixels)."""
    dim = next(d for d, n in enumerate(a.shape) if n == 3)
    return torch.linalg.cross(a, b, dim=dim)


def ray_condition(K, c2w, H, W, device, flip_flag=None):
    """K [B, V, 4] (fx, fy, cx, cy in pixels), c2w [B, V, 4, 4] -> Pluecker coordinates (o x d, d) per pixel, [B, V, H, W, 6]."""
    B, V = K.shape[:2]
    rows, cols = custom_meshgrid(torch.linspace(0, H - 1, H, device=device, dtype=c2w.dtype),
                                 torch.linspace(0, W - 1, W, device=device, dtype=c2w.dtype))
    u = cols.reshape(1, 1, H * W).expand(B, V, H * W) + 0.5
    v = rows.reshape(1, 1, H * W).expand(B, V, H * W) + 0.5
    if flip_flag is not None and int(torch.sum(flip_flag)) > 0:
        rows_f, cols_f = custom_meshgrid(torch.linspace(0, H - 1, H, device=device, dtype=c2w.dtype),
                                         torch.linspace(W - 1, 0, W, device=device, dtype=c2w.dtype))
        u[:, flip_flag, ...] = cols_f.reshape(1, 1, H * W).expand(B, 1, H * W) + 0.5
        v[:, flip_flag, ...] = rows_f.reshape(1, 1, H * W).expand(B, 1, H * W) + 0.5
    fx, fy, cx, cy = K.chunk(4, dim=-1)
    ones = torch.ones_like(u)
    x = (u - cx) / fx * ones
    y = (v - cy) / fy * ones
    d = torch.stack((x, y, ones.expand_as(y)), dim=-1)
    d = d / d.norm(dim=-1, keepdim=True)
    rays_d = d @ c2w[..., :3, :3].transpose(-1, -2)
    rays_o = c2w[..., :3, 3][:, :, None].expand_as(rays_d)
    return torch.cat([_legacy_cross(rays_o, rays_d), rays_d], dim=-1).reshape(B, c2w.shape[1], H, W, 6)


class RealEstate10KPoseProcessor:
    def __init__(self, sample_stride=4, minimum_sample_stride=1, sample_n_frames=16, relative_pose=False, zero_t_first_frame=False,
                 sample_size=[256, 384], rescale_fxy=False, shuffle_frames=False, use_flip=False, return_clip_name=False, is_i2v=False):
        self.relative_pose, self.zero_t_first_frame = relative_pose, zero_t_first_frame
        self.sample_stride, self.minimum_sample_stride, self.sample_n_frames = sample_stride, minimum_sample_stride, sample_n_frames
        self.return_clip_name, self.is_i2v = return_clip_name, is_i2v
        self.sample_size = (sample_size, sample_size) if isinstance(sample_size, int) else tuple(sample_size)
        self.rescale_fxy, self.shuffle_frames, self.use_flip = rescale_fxy, shuffle_frames, use_flip
        self.sample_wh_ratio = self.sample_size[1] / self.sample_size[0]
        # kept for attribute parity: [resize, (flip), normalise]; only the flip module is used here (its flag)
        self.pixel_transforms = [None, RandomHorizontalFlipWithPose(), None] if use_flip else [None, None]

    def get_relative_pose(self, cam_params):
        """Poses relative to the first camera, which is placed at the origin (or at its original distance below it on -y)."""
        first_c2w, first_w2c = cam_params[0].c2w_mat, cam_params[0].w2c_mat
        lift = 0 if self.zero_t_first_frame else np.linalg.norm(first_c2w[:3, 3])
        anchor = np.array([[1, 0, 0, 0], [0, 1, 0, -lift], [0, 0, 1, 0], [0, 0, 0, 1]])
        to_rel = anchor @ first_w2c
        return np.array([anchor] + [to_rel @ cam.c2w_mat for cam in cam_params[1:]], dtype=np.float32)

    def load_cameras(self, pose_file):
        with open(pose_file, "r") as f:
            lines = f.readlines()
        if "youtube" in lines[0]:
            lines = lines[1:]
        return [Camera([float(x) for x in line.strip().split(" ")]) for line in lines]

    def _embed(self, cams, image_path=None):
        n = self.sample_n_frames
        assert len(cams) >= n
        stride = self.sample_stride
        if len(cams) < n * stride:
            stride = random.randint(self.minimum_sample_stride, int(len(cams) // n))
        end = min(n * stride, len(cams))
        assert end >= n
        picks = np.linspace(0, end - 1, n, dtype=int)
        if self.shuffle_frames:
            picks = picks[np.random.permutation(n)]
        cams = [cams[i] for i in picks]
        Hs, Ws = self.sample_size
        if self.rescale_fxy:
            from PIL import Image
            ow, oh = Image.open(image_path).size
            if ow / oh > self.sample_wh_ratio:
                for c in cams:
                    c.fx = Hs * (ow / oh) * c.fx / Ws
            else:
                for c in cams:
                    c.fy = Ws / (ow / oh) * c.fy / Hs
        K = torch.as_tensor(np.asarray([[c.fx * Ws, c.fy * Hs, c.cx * Ws, c.cy * Hs] for c in cams], dtype=np.float32))[None]
        poses = self.get_relative_pose(cams) if self.relative_pose else np.array([c.c2w_mat for c in cams], dtype=np.float32)
        c2w = torch.as_tensor(poses)[None]
        flip = self.pixel_transforms[1].get_flip_flag(n) if self.use_flip else torch.zeros(n, dtype=torch.bool, device=c2w.device)
        return ray_condition(K, c2w, Hs, Ws, device="cpu", flip_flag=flip)

    def get_plucker_embedding(self, pose_file, image_path=None):
        """RealEstate10K pose file -> [1, n_frames, H, W, 6]."""
        return self._embed(self.load_cameras(pose_file), image_path)

    def get_plucker_embedding_direct_from_cam_params(self, pose_enc, image_size, image_path=None):
        """VGGT pose encoding [1, n, 9] (absT_quaR_FoV) -> [1, n_frames, H, W, 6] (the CLI path)."""
        extrinsic, intrinsic = pose_encoding_to_extri_intri(pose_enc, image_size, pose_encoding_type="absT_quaR_FoV")
        cams = create_camera_params_from_batch(extrinsic.cpu().numpy().squeeze(0), intrinsic.cpu().numpy().squeeze(0))
        return self._embed(cams, image_path)
