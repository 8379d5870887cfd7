from .models.wan_video_dit import WanModel
from .pipelines.wan_video_new import ModelConfig, WanVideoPipeline
from .schedulers.flow_match import FlowMatchScheduler
