"""Property-based checks (hypothesis) of the host-side planners — pure arithmetic, no device:
the attention tile-schedule planner behind fwb_attn_fwd (fwb_attn_plan), the VAE tile lists / blending masks (tiled decode and encode),
the frame-aligned shard layout of the sequence-parallel context, and the pipeline's shape rounding."""
import ctypes as C

import torch
from hypothesis import given, settings, strategies as st


@settings(max_examples=300, deadline=None, derandomize=True)
@given(B=st.integers(1, 24), H=st.integers(1, 48), Lq=st.integers(1, 40000), Lk=st.integers(1, 40000), D=st.sampled_from([64, 96, 128]),
       sms=st.sampled_from([74, 132, 148, 160]), ws_mb=st.sampled_from([0, 1, 16, 78, 512]))
def test_attention_plan_invariants(B, H, Lq, Lk, D, sms, ws_mb):
    import fwb200
    n_full, S = C.c_int(), C.c_int()
    ws = ws_mb << 20
    assert fwb200.lib.fwb_attn_plan(B, H, Lq, Lk, D, ws, sms, C.byref(n_full), C.byref(S)) == 0
    n_tiles = -(-Lq // 256) * H * B
    assert 0 <= n_full.value <= n_tiles and 1 <= S.value <= 16
    tail = n_tiles - n_full.value
    if S.value == 1:
        assert tail == 0                                   # nothing is left for a split that does not happen
    else:
        assert n_full.value % sms == 0 and 0 < tail < sms  # only the partly filled last wave is split ...
        assert tail * S.value * 256 * (D + 1) * 4 <= ws    # ... its fp32 partials + lse fit the workspace ...
        assert Lk // S.value >= 512                        # ... and every split keeps enough keys to be worth a CTA
        # ... and the modelled time (rounds of 1/S-length CTAs + a fixed merge cost) beats the unsplit schedule by at least 7 %
        waves = n_full.value // sms
        split_cost = waves + (-(-tail * S.value // sms)) / S.value + 0.04
        assert split_cost <= 0.93 * (waves + 1) + 1e-9


@settings(max_examples=200, deadline=None, derandomize=True)
@given(H=st.integers(1, 70), W=st.integers(1, 110), sh=st.integers(1, 40), sw=st.integers(1, 60), dh=st.integers(1, 40), dw=st.integers(1, 60))
def test_vae_tiles_cover_the_grid_with_positive_weight(H, W, sh, sw, dh, dw):
    """Every latent cell is covered by at least one tile and the accumulated blending weight is positive everywhere (no 0 / 0 in
    `values / weight`), for any tile size > stride — the reference's rule, inference uses (30, 52) / (15, 26)."""
    from FantasyWorld.diffsynth_wan21.models.wan_video_vae import WanVideoVAE
    size, stride = (max(sh, dh) + 1, max(sw, dw) + 1), (min(sh, dh), min(sw, dw))         # overlapping tiles: size > stride (as the reference needs)
    tasks = WanVideoVAE.tile_tasks(H, W, size, stride)
    weight = torch.zeros(H, W)
    vae = WanVideoVAE.__new__(WanVideoVAE)                 # build_mask needs no parameters
    border = (size[0] - stride[0], size[1] - stride[1])
    for h0, h1, w0, w1 in tasks:
        th, tw = min(h1, H) - h0, min(w1, W) - w0
        assert th > 0 and tw > 0
        mask = WanVideoVAE.build_mask(vae, torch.zeros(1, 1, 1, th, tw), (h0 == 0, h1 >= H, w0 == 0, w1 >= W), border)[0, 0, 0]
        assert float(mask.min()) > 0 and float(mask.max()) <= 1
        weight[h0:h0 + th, w0:w0 + tw] += mask
    assert float(weight.min()) > 0
    assert tasks[0][0] == 0 and tasks[0][2] == 0 and len(set(tasks)) == len(tasks)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(h=st.integers(1, 2000), w=st.integers(1, 2000), f=st.integers(1, 400))
def test_wan22_shape_rounding(h, w, f):
    from FantasyWorld.diffsynth_wan22.pipelines.wan_video_new import WanVideoPipeline
    pipe = WanVideoPipeline(device="cpu")
    H, W, F = pipe.check_resize_height_width(h, w, f)
    assert H % 16 == 0 and W % 16 == 0 and F % 4 == 1 and 0 <= H - h < 16 and 0 <= W - w < 16 and 0 <= F - f <= 4
    assert pipe.check_resize_height_width(H, W, F) == (H, W, F)          # idempotent


@settings(max_examples=200, deadline=None, derandomize=True)
@given(world=st.integers(1, 8), f=st.integers(1, 40), h=st.integers(1, 48), w=st.integers(1, 80))
def test_sequence_parallel_layout_partitions_both_streams(world, f, h, w):
    """Video rows: contiguous shards differing by at most one row; geometry rows: whole frames per rank (frame attention needs no
    exchange), differing by at most one frame; both cover their stream exactly once, in rank order."""
    from fwb200.sp import SPLayout
    lay = SPLayout(world=world, f=f, h=h, w=w)
    for ranges, total, sizes in ((([lay.video_range(r) for r in range(world)]), lay.L, lay.video_rows),
                                (([lay.geo_range(r) for r in range(world)]), lay.N, lay.geo_rows())):
        assert ranges[0][0] == 0 and ranges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert [b - a for a, b in ranges] == list(sizes) and sum(sizes) == total
    assert max(lay.video_rows) - min(lay.video_rows) <= 1 and max(lay.frames) - min(lay.frames) <= 1
    assert all(r % lay.P == 0 for r in lay.geo_rows()) and lay.P == 5 + h * w
    assert sorted(lay.video_rows, reverse=True) == lay.video_rows          # the larger shards come first


@settings(max_examples=100, deadline=None, derandomize=True)
@given(n=st.integers(1, 200), shift=st.floats(1.0, 12.0), extra=st.booleans())
def test_flow_match_schedule_properties(n, shift, extra):
    """Sigmas start at 1 (shift maps 1 -> 1), decrease strictly, every Euler increment is negative and the increments telescope to
    -sigma_0 (the sampler integrates the whole way to sigma = 0); a constant velocity field is integrated exactly."""
    from FantasyWorld.diffsynth_wan21.schedulers.flow_match import FlowMatchScheduler
    s = FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=extra)
    s.set_timesteps(n)
    sig = s.sigmas
    assert len(sig) == n and abs(float(sig[0]) - 1.0) < 1e-6 and bool((sig[1:] < sig[:-1]).all()) and float(sig[-1]) >= 0.0
    assert torch.equal(s.timesteps, sig * 1000)
    steps = [s.dsigma(t) for t in s.timesteps]
    assert all(d < 0 for d in steps[:-1]) and steps[-1] <= 0
    assert abs(sum(steps) + float(sig[0])) < 1e-4
    x = torch.zeros(3)
    for t in s.timesteps:
        x = s.step(torch.ones(3), t, x)
    assert torch.allclose(x, torch.full((3,), -float(sig[0])), atol=1e-4)


def test_bench_flop_accounting_matches_baseline_table():
    """bench.forward_flops is the denominator of every throughput / roofline figure.  It must agree with BASELINE.md §2 (2.134 PFLOP per C2
    forward, 8.548 per C4 forward, 0.758 at the Wan2.1 CLI default, 21.70 for the 81-latent-frame stress row) and never exceed it: the
    bench leaves out what the engine hoists out of the loop (context K/V projections, camera group-1 MLP, embeddings), which is at most
    1 % of a forward, so the reported TFLOP/s can only be conservative.  Per-call attention figures: 21.98 / 9.92 / 4.42 TFLOP."""
    import bench
    rows = (((21, 30, 52, 16, 24), {}, 2.134), ((21, 45, 80, 16, 24), dict(clip=False, camera_adaln=False), 8.548),
            ((21, 21, 37, 16, 24), {}, 0.758), ((81, 30, 52, 16, 24), {}, 21.70))
    for args, kw, ref in rows:
        ours = bench.forward_flops(*args, **kw) / 1e15
        assert 0.988 * ref <= ours <= 1.0005 * ref, (args, ours, ref)
    L, N = 21 * 30 * 52, 21 * (5 + 30 * 52)
    assert round(4 * 40 * L * L * 128 / 1e12, 2) == 21.98 and round(2 * 4 * 12 * L * N * 96 / 1e12, 2) == 9.92
    assert round(4 * 16 * N * N * 64 / 1e12, 2) == 4.42
    assert bench.forward_flops(21, 30, 52, 16, 24) < bench.forward_flops(22, 30, 52, 16, 24) < bench.forward_flops(22, 31, 52, 16, 25)
