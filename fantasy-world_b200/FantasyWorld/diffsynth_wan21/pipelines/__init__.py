from .wan_video import WanVideoPipeline
