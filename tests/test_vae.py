"""Wan VAE mirror (tiled decode after the sampler loop; N3: first-frame encode) against goldens written by the
UNMODIFIED reference (tools/make_golden_vae.py): same state_dict schema, whole-clip evaluation == the reference's frame-by-frame
streaming with caches, tiled blending in the reference's order, and tile sharding over a gloo group."""
import os
import socket

import pytest
import torch

from _common import gold, rel_err


def _vae():
    from FantasyWorld.diffsynth_wan21.models.wan_video_vae import WanVideoVAE
    from fwb_synth import synth_init
    wrap = torch.nn.Module()
    wrap.vae = WanVideoVAE(z_dim=16)
    wrap.vae.model.requires_grad_(True)
    synth_init(wrap, seed=0, gen_device="cpu")
    wrap.vae.model.requires_grad_(False)
    return wrap, wrap.vae.eval()


def _inputs():
    g = torch.Generator().manual_seed(11)
    return dict(z=torch.randn(1, 16, 3, 6, 8, generator=g), video=torch.randn(3, 9, 32, 48, generator=g).clamp(-1, 1))


def test_vae_schema_decode_encode_match_reference():
    g = gold("vae.pt")
    wrap, vae = _vae()
    assert {k: list(v.shape) for k, v in wrap.state_dict().items()} == g["schema"]          # 194 keys: the released VAE checkpoint loads
    inp = _inputs()
    with torch.no_grad():
        single = vae.decode(inp["z"], device="cpu", tiled=False)
        tiled = vae.decode(inp["z"], device="cpu", tiled=True, tile_size=g["tile_size"], tile_stride=g["tile_stride"])
        enc = vae.encode([inp["video"]], device="cpu")
    assert single.shape == g["single"].shape == (1, 3, 9, 48, 64)
    # fp32, whole-clip convolutions vs 3 streamed single-frame passes: identical math, cuDNN/oneDNN summation order differs
    assert rel_err(single, g["single"]) < 2e-5, rel_err(single, g["single"])
    assert rel_err(tiled, g["tiled"]) < 2e-5, rel_err(tiled, g["tiled"])
    assert rel_err(enc, g["enc"]) < 2e-5, rel_err(enc, g["enc"])
    assert rel_err(g["tiled"], g["single"]) > 1e-3          # tiling really changes the result (limited receptive field), so the test is sharp


def test_tile_task_list_matches_reference_rule():
    from FantasyWorld.diffsynth_wan21.models.wan_video_vae import WanVideoVAE
    # inference_wan21.py:324-330: 60x104 latent, tile (30, 52), stride (15, 26) -> 3 x 3 tiles, the last ones clipped by slicing
    t = WanVideoVAE.tile_tasks(60, 104, (30, 52), (15, 26))
    assert t == [(h, h + 30, w, w + 52) for h in (0, 15, 30) for w in (0, 26, 52)]
    assert WanVideoVAE.tile_tasks(6, 8, (4, 6), (2, 3)) == [(0, 4, 0, 6), (0, 4, 3, 9), (2, 6, 0, 6), (2, 6, 3, 9)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        g = gold("vae.pt")
        _, vae = _vae()
        with torch.no_grad():
            out = vae.decode(_inputs()["z"], device="cpu", tiled=True, tile_size=g["tile_size"], tile_stride=g["tile_stride"], group=dist.group.WORLD)
        ret[rank] = rel_err(out, g["tiled"])
    finally:
        dist.destroy_process_group()


def test_tiled_decode_sharded_over_two_ranks_gloo():
    """Tiles dealt round-robin to 2 ranks, accumulators all-reduced: same video on every rank (sum order differs: fp32 rounding)."""
    import torch.multiprocessing as mp
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] < 2e-5 and ret[1] < 2e-5, dict(ret)
