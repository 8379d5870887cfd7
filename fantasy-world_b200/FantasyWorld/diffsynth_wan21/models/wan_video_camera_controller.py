"""Mirror of the hot-path part of FantasyWorld/diffsynth_wan2{1,2}/models/wan_video_camera_controller.py: `SimpleAdapter`,
the Wan2.2-Fun control adapter that turns the camera latents [b, 24, f, H, W] (Plücker rays, 4 video frames folded into
channels) into a [b, dim, f, H/16, W/16] tensor added to the patch embedding (ref: wan_video_camera_controller.py:8-76,
diffsynth_wan22/models/wan_video_dit.py:390-396).

It depends only on the camera path, not on the latent or the timestep, so the fusion core evaluates it ONCE per sample and
folds the result into the patchify GEMM epilogue as a residual (FusionCore._control_tokens); the reference re-runs its two
5120-channel 3x3 convolutions on every forward (71 TFLOP at 720p).  The convolutions themselves are
torch/cuDNN calls — once per sample, outside the denoising loop.  The camera-trajectory helpers of the reference file
(generate_camera_coordinates, process_pose_file, ray_condition) are CPU pre-processing and out of scope.
"""
import torch
import torch.nn as nn


class ResidualBlock(nn.Module):
    """x + conv2(relu(conv1(x))).  ref: wan_video_camera_controller.py:64-76."""

    def __init__(self, dim):
        super().__init__()
        self.conv1 = nn.Conv2d(dim, dim, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(dim, dim, kernel_size=3, padding=1)

    def forward(self, x):
        return x + self.conv2(torch.relu(self.conv1(x)))


class SimpleAdapter(nn.Module):
    def __init__(self, in_dim, out_dim, kernel_size, stride, num_residual_blocks=1):
        super().__init__()
        self.pixel_unshuffle = nn.PixelUnshuffle(downscale_factor=8)
        self.conv = nn.Conv2d(in_dim * 64, out_dim, kernel_size=kernel_size, stride=stride, padding=0)
        self.residual_blocks = nn.Sequential(*[ResidualBlock(out_dim) for _ in range(num_residual_blocks)])

    def forward(self, x):
        """[bs, c, f, h, w] -> [bs, out_dim, f, h/16, w/16]: frames are folded into the batch for the 2-D convolutions."""
        bs, c, f, h, w = x.shape
        y = self.pixel_unshuffle(x.permute(0, 2, 1, 3, 4).reshape(bs * f, c, h, w))
        y = self.residual_blocks(self.conv(y))
        return y.view(bs, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)
