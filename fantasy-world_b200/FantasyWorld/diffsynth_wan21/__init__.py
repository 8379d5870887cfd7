from .models import ModelManager
from .pipelines import WanVideoPipeline
from .schedulers import FlowMatchScheduler
