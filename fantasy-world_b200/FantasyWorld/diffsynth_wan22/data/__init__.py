"""Camera-pose pre-processing used by inference_wan22.py (same module as the Wan2.1 tree: the reference's two copies are identical)."""
