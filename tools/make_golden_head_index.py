"""Golden for the geometry heads' INDEX behaviour ("index selection / chunking must be bit-identical").

Runs the UNMODIFIED reference heads (DPTHead_3D_Causal, CameraHead; CPU fp32) on integer-coded token lists and records, through
forward pre-hooks on sub-modules that exist under the same names in this repo's mirror, WHICH tokens reach WHICH stage in WHICH
order: the layer selection [23,17,11,7], the `[:, f0:f1, patch_start_idx:]` slices of the 4-latent-frame chunks, the
16-video-frame chunks of the fusion stage, the camera-token slice and its 4x temporal expansion.

    python tools/make_golden_head_index.py      # build container only -> tests/golden/head_index.pt  (a few kB of int16)
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "fantasy-world_b200"))

sys.path.insert(0, str(ROOT / "tests"))
from _head_index import S, GH, GW, record_head_indexing  # noqa: E402


def main():
    from ref_shim import import_reference
    from fwb_synth import synth_init
    import importlib
    import_reference()
    dpt = importlib.import_module("FantasyWorld.vggt.heads.dpt_head")
    cam = importlib.import_module("FantasyWorld.vggt.heads.camera_head")
    torch.manual_seed(0)
    wrap = torch.nn.Module()
    wrap.vggt = torch.nn.Module()
    wrap.vggt.depth_head = dpt.DPTHead_3D_Causal(dim_in=2048, output_dim=2, activation="exp", conf_activation="expp1", patch_size=16)
    wrap.vggt.camera_head = cam.CameraHead(dim_in=2048)
    synth_init(wrap, seed=0, gen_device="cpu")
    wrap.eval()
    rec = record_head_indexing(wrap.vggt.depth_head, wrap.vggt.camera_head)
    out = ROOT / "tests" / "golden" / "head_index.pt"
    torch.save({"S": S, "GH": GH, "GW": GW, "records": rec}, out)
    for name, shape, val in rec:
        print(name, shape, None if val is None else tuple(val.shape))
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
