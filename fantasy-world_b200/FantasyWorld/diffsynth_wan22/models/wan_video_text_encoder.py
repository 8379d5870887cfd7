"""Wan2.2 location of the umT5 encoder (the reference's two copies are byte-identical: same module as the Wan2.1 tree)."""
from ...diffsynth_wan21.models.wan_video_text_encoder import *  # noqa: F401,F403
from ...diffsynth_wan21.models.wan_video_text_encoder import (GELU, T5Attention, T5FeedForward, T5LayerNorm,  # noqa: F401
                                                              T5RelativeEmbedding, T5SelfAttention, WanTextEncoder,
                                                              WanTextEncoderStateDictConverter, fp16_clamp, init_weights)
