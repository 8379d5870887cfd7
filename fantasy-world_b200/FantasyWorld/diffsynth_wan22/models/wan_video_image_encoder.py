"""Wan2.2 location of the CLIP image tower (the reference's two copies are byte-identical: same module as the Wan2.1 tree)."""
from ...diffsynth_wan21.models.wan_video_image_encoder import *  # noqa: F401,F403
from ...diffsynth_wan21.models.wan_video_image_encoder import (AttentionBlock, AttentionPool, LayerNorm, QuickGELU,  # noqa: F401
                                                               SelfAttention, SwiGLU, VisionTransformer, WanImageEncoder,
                                                               WanImageEncoderStateDictConverter, XLMRobertaCLIP,
                                                               clip_xlm_roberta_vit_h_14, pos_interpolate)
