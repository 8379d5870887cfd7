"""Golden for the camera pre-processing mirror (FantasyWorld/diffsynth_wan21/data/dataset_re10k.py): the UNMODIFIED reference
`RealEstate10KPoseProcessor` on small clips — the CLI configuration (inference_wan21.py:172-182), the other pose conventions, a
RealEstate10K pose file, the 3-frame case (where the reference's dim-less `torch.cross` picks the frame axis), flips and fx/fy rescaling.

    python tools/make_golden_pose_processor.py      # build container -> tests/golden/pose_processor.pt (~25 KB)
"""
from __future__ import annotations

import importlib
import random
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))


def trajectory(n, seed):
    """n world-to-camera [3, 4] matrices (smooth random rotation + translation) and pixel intrinsics for a 6 x 8 image."""
    g = torch.Generator().manual_seed(seed)
    ext, intr = [], []
    for i in range(n):
        w = torch.randn(3, generator=g) * 0.15 * (i + 1)
        W = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        R = torch.linalg.matrix_exp(W)
        t = torch.randn(3, 1, generator=g) * 0.3 + torch.tensor([[0.1 * i], [0.5], [1.0]])
        ext.append(torch.cat([R, t], dim=1))
        intr.append(torch.tensor([[7.5, 0, 4.0], [0, 7.0, 3.0], [0, 0, 1.0]]))
    return torch.stack(ext)[None], torch.stack(intr)[None]


def pose_file_text(n, seed):
    ext, _ = trajectory(n, seed)
    lines = ["https://www.youtube.com/watch?v=golden"]
    for i, e in enumerate(ext[0]):
        lines.append(" ".join([str(1000 * i), "0.52", "0.93", "0.5", "0.5", "0", "0"] + [repr(float(x)) for x in e.flatten()]))
    return "\n".join(lines) + "\n"


CASES = {
    "cli": dict(sample_stride=1, sample_n_frames=5, relative_pose=True, zero_t_first_frame=True, sample_size=[6, 8], rescale_fxy=False,
                shuffle_frames=False, use_flip=False, is_i2v=True),
    "relative_lifted": dict(sample_stride=1, sample_n_frames=5, relative_pose=True, zero_t_first_frame=False, sample_size=[6, 8]),
    "absolute": dict(sample_stride=2, sample_n_frames=4, relative_pose=False, sample_size=[6, 8]),
    "three_frames": dict(sample_stride=1, sample_n_frames=3, relative_pose=True, zero_t_first_frame=True, sample_size=[6, 8]),
    "flip": dict(sample_stride=1, sample_n_frames=5, relative_pose=True, sample_size=[6, 8], use_flip=True),
    "short_clip": dict(sample_stride=4, minimum_sample_stride=1, sample_n_frames=4, relative_pose=True, sample_size=[6, 8]),
}


def main():
    from ref_shim import import_reference
    import_reference()
    ds = importlib.import_module("FantasyWorld.diffsynth_wan21.data.dataset_re10k")
    pe = importlib.import_module("FantasyWorld.vggt.utils.pose_enc")
    ext, intr = trajectory(9, seed=1)
    pose_enc = pe.extri_intri_to_pose_encoding(ext, intr, [6, 8], pose_encoding_type="absT_quaR_FoV")
    out = {"pose_enc": pose_enc, "cases": CASES, "direct": {}, "file": {}}
    text = pose_file_text(9, seed=2)
    out["pose_file"] = text
    with tempfile.TemporaryDirectory() as td:
        pf = Path(td) / "clip.txt"
        pf.write_text(text)
        from PIL import Image
        img = Path(td) / "frame.png"
        Image.new("RGB", (20, 6)).save(img)          # wider than 8:6 -> fx is rescaled
        for name, kw in CASES.items():
            for tag, fn in (("direct", lambda p: p.get_plucker_embedding_direct_from_cam_params(pose_enc, image_size=(6, 8))),
                            ("file", lambda p: p.get_plucker_embedding(str(pf)))):
                torch.manual_seed(4)
                random.seed(4)
                np.random.seed(4)
                out[tag][name] = fn(ds.RealEstate10KPoseProcessor(**kw))
        p = ds.RealEstate10KPoseProcessor(sample_stride=1, sample_n_frames=5, relative_pose=True, sample_size=[6, 8], rescale_fxy=True)
        out["rescale_file"] = p.get_plucker_embedding(str(pf), image_path=str(img))
    path = ROOT / "tests" / "golden" / "pose_processor.pt"
    torch.save(out, path)
    print("wrote", path, path.stat().st_size, "bytes", {k: tuple(v.shape) for k, v in out["direct"].items()})


if __name__ == "__main__":
    main()
