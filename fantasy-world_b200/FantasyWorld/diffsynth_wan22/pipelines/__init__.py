from .wan_video_new import ModelConfig, WanVideoPipeline
