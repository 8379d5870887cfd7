"""Wan2.2 location of the control adapter (same module as the Wan2.1 tree)."""
from ...diffsynth_wan21.models.wan_video_camera_controller import ResidualBlock, SimpleAdapter  # noqa: F401
