"""Wan2.2 location of the Wan VAE.  The 16-channel `WanVideoVAE` that Wan2.2-Fun-A14B uses (`Wan2.1_VAE.pth`, model_wan22.py:158-162) is
the same network as in the Wan2.1 tree — the golden of tests/test_wan22_condition.py is written by THIS tree's reference class and
checked against the shared mirror.  The 48-channel `WanVideoVAE38` of Wan2.2-TI2V-5B (wan_video_vae.py:1278-1391 in the reference) is
not used by FantasyWorld and is not mirrored."""
from ...diffsynth_wan21.models.wan_video_vae import (AttentionBlock, CausalConv3d, Decoder3d, Encoder3d, Resample,  # noqa: F401
                                                     ResidualBlock, RMS_norm, Upsample, VideoVAE_, WanVideoVAE,
                                                     WanVideoVAEStateDictConverter)


class WanVideoVAE38(WanVideoVAE):
    def __init__(self, *a, **k):
        raise NotImplementedError("WanVideoVAE38 (Wan2.2-TI2V-5B) is outside the FantasyWorld path")
