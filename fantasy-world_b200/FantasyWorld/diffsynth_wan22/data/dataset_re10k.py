"""Wan2.2 location of the pose processor (the reference's two copies are byte-identical: same module as the Wan2.1 tree)."""
from ...diffsynth_wan21.data.dataset_re10k import (Camera, RandomHorizontalFlipWithPose, RealEstate10KPoseProcessor,  # noqa: F401
                                                   create_camera_params_from_batch, custom_meshgrid, ray_condition)
