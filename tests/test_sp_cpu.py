"""Sequence-parallel host logic on CPU: world_size 2, 3 and 8 gloo process groups.  Checks the shard layout (video rows,
frame-aligned geometry rows), the packed all-gather (equal and ragged shards) and that "local queries x gathered keys"
reproduces the rows of the un-sharded attention (oracle math) — the only cross-rank dependency of the hot path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fwb200.sp import SPContext
        from oracle import fw_oracle as O
        sp = SPContext()
        f, h, w = 5, 2, 3
        lay = sp.set_grid(f, h, w)
        assert sum(lay.video_rows) == f * h * w and sum(lay.frames) == f
        assert lay.geo_rows() == [fr * (5 + h * w) for fr in lay.frames]
        torch.manual_seed(0)                       # same tensors on every rank
        L, N, H, D = lay.L, lay.N, 2, 8
        q = torch.randn(1, H, L, D)
        k = torch.randn(1, H, N, D)
        v = torch.randn(1, H, N, D)
        full = O.sdpa(q, k, v)                      # video queries x geometry keys (adapter direction 1)
        # each rank owns geometry rows g0:g1 (ragged across ranks) and video rows r0:r1
        g0, g1 = lay.geo_range(rank)
        r0, r1 = lay.video_range(rank)
        kv_local = torch.cat([k[0].transpose(0, 1)[g0:g1].reshape(g1 - g0, H * D), v[0].transpose(0, 1)[g0:g1].reshape(g1 - g0, H * D)], dim=1)
        kv_all = sp.all_gather_rows(kv_local.contiguous(), lay.geo_rows())
        assert kv_all.shape == (N, 2 * H * D)
        k_all = kv_all[:, : H * D].view(N, H, D).transpose(0, 1)[None]
        v_all = kv_all[:, H * D:].view(N, H, D).transpose(0, 1)[None]
        assert torch.equal(k_all, k) and torch.equal(v_all, v)
        mine = O.sdpa(q[:, :, r0:r1], k_all, v_all)
        assert torch.allclose(mine, full[:, :, r0:r1], atol=1e-6)
        # equal-size gather (video rows when L % world == 0, else ragged too)
        x_local = torch.arange(r0, r1, dtype=torch.float32)[:, None].repeat(1, 3)
        x_all = sp.all_gather_rows(x_local.contiguous(), lay.video_rows)
        assert torch.equal(x_all[:, 0], torch.arange(L, dtype=torch.float32))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sp_layout_and_gather_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_layout_at_baseline_sizes():
    from fwb200.sp import SPLayout
    lay = SPLayout(8, 21, 30, 52)
    assert lay.video_rows == [4095] * 8 and lay.frames == [3, 3, 3, 3, 3, 2, 2, 2]
    assert lay.geo_range(7) == ((21 - 2) * 1565, 21 * 1565) and lay.video_range(3) == (3 * 4095, 4 * 4095)
    lay = SPLayout(4, 21, 45, 80)
    assert lay.video_rows == [18900] * 4 and lay.frames == [6, 5, 5, 5]


def _cfg_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fwb200.sp import CFGParallel
        cp = CFGParallel()
        half = world // 2
        assert cp.role == (0 if rank < half else 1) and cp.half == half
        assert (cp.sp is None) == (half == 1)
        if cp.sp is not None:                                # each half is its own sequence-parallel group
            assert cp.sp.world == half and cp.sp.rank == rank % half
            lay = cp.sp.set_grid(5, 2, 3)
            r0, r1 = lay.video_range(cp.sp.rank)
            rows = torch.arange(r0, r1, dtype=torch.float32)[:, None] + 1000.0 * cp.role
            allr = cp.sp.all_gather_rows(rows.contiguous(), lay.video_rows)
            assert torch.equal(allr[:, 0], torch.arange(lay.L, dtype=torch.float32) + 1000.0 * cp.role)   # never mixes the halves
        # every rank of a half holds that half's prediction; the pair swap gives (conditional, unconditional) everywhere
        mine = torch.full((1, 4, 2, 3), float(10 + cp.role))
        pos, neg = cp.exchange(mine)
        assert torch.equal(pos, torch.full_like(mine, 10.0)) and torch.equal(neg, torch.full_like(mine, 11.0))
        assert cp.n_exchanges == 1 and cp.exchange_bytes == 2 * mine.numel() * 4
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cfg_parallel_groups_and_exchange_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cfg_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))
