from .model_manager import ModelManager
