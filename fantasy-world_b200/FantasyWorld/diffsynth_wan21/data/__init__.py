"""Camera-pose pre-processing used by the inference CLIs (`dataset_re10k`).  The reference's video I/O helpers (`video.py`: imageio
readers / writers) are not part of this build."""
