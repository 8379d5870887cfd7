"""Camera pre-processing mirror (FantasyWorld/diffsynth_wan21|22/data/dataset_re10k.py) against goldens written by the UNMODIFIED
reference (tools/make_golden_pose_processor.py): Pluecker embeddings must be EQUAL bit for bit — it is index / float32 arithmetic in a
fixed order — for the CLI configuration, the other pose conventions, a RealEstate10K pose file, the 3-frame corner case of the
reference's dim-less torch.cross, random flips (same RNG draws) and fx / fy rescaling."""
import random

import numpy as np
import pytest
import torch

from _common import gold


@pytest.mark.parametrize("tree", ["diffsynth_wan21", "diffsynth_wan22"])
def test_plucker_embeddings_bit_equal_to_reference(tree, tmp_path):
    import importlib
    from PIL import Image
    ds = importlib.import_module(f"FantasyWorld.{tree}.data.dataset_re10k")
    g = gold("pose_processor.pt")
    pf = tmp_path / "clip.txt"
    pf.write_text(g["pose_file"])
    for name, kw in g["cases"].items():
        for tag, fn in (("direct", lambda p: p.get_plucker_embedding_direct_from_cam_params(g["pose_enc"], image_size=(6, 8))),
                        ("file", lambda p: p.get_plucker_embedding(str(pf)))):
            torch.manual_seed(4)
            random.seed(4)
            np.random.seed(4)
            out = fn(ds.RealEstate10KPoseProcessor(**kw))
            ref = g[tag][name]
            assert out.dtype == ref.dtype == torch.float32 and out.shape == ref.shape
            assert torch.equal(out, ref), (tree, tag, name, float((out - ref).abs().max()))
    img = tmp_path / "frame.png"
    Image.new("RGB", (20, 6)).save(img)
    p = ds.RealEstate10KPoseProcessor(sample_stride=1, sample_n_frames=5, relative_pose=True, sample_size=[6, 8], rescale_fxy=True)
    assert torch.equal(p.get_plucker_embedding(str(pf), image_path=str(img)), g["rescale_file"])
    # sanity of the geometry itself: unit ray directions, moment orthogonal to the direction
    e = g["direct"]["cli"]
    d, m = e[..., 3:], e[..., :3]
    assert torch.allclose(d.norm(dim=-1), torch.ones(1, 5, 6, 8), atol=1e-6) and float((d * m).sum(-1).abs().max()) < 1e-5
    assert float(m[0, 0].abs().max()) == 0.0          # first camera at the origin (zero_t_first_frame): zero moments


def test_pose_processor_feeds_the_pose_encoder_shape():
    """The CLI hands [1, 81, H, W, 6] to generate_video; here: the same layout at a small size."""
    from FantasyWorld.diffsynth_wan21.data.dataset_re10k import RealEstate10KPoseProcessor
    g = gold("pose_processor.pt")
    p = RealEstate10KPoseProcessor(sample_stride=1, sample_n_frames=9, relative_pose=True, zero_t_first_frame=True, sample_size=[16, 32],
                                   is_i2v=True)
    out = p.get_plucker_embedding_direct_from_cam_params(g["pose_enc"], image_size=(16, 32))
    assert out.shape == (1, 9, 16, 32, 6) and torch.isfinite(out).all()
