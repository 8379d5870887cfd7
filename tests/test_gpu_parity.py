"""GPU parity of the hot path, end to end through the reference-shaped module API (FantasyWorld.*) and the C-ABI kernels:

  CUDA path (bf16)   vs   golden vectors written by the UNMODIFIED reference (CPU fp32, tools/make_golden.py)
                     vs   the oracle with bf16 rounding emulated at the reference's autocast rounding points.

Protocol: the north-star tolerance rtol=1e-3/atol=1e-4 is below one bf16 ulp (2^-8), so for bf16 tensors
we assert (i) closeness to the bf16-emulating oracle at a few bf16 ulps, and (ii) that the error against the fp32 golden
is bf16-sized; fp32-output kernels are held to rtol=1e-3/atol=1e-4 in test_gpu_ops.py.  Index paths are bit-exact.
"""
import pytest
import torch

from _common import gold, max_err, rel_err, synth_state_dict

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fp32_rope_angles():
    """The fp32 goldens come from the reference run WITHOUT autocast (fp32 VGGT RoPE angles); the product default follows the
    reference under autocast (bf16 angles, see fwb200.engine.rope2d_expanded).  Tests against the fp32 goldens switch modes."""
    import fwb200.engine as E
    E.ROPE2D_FP32_ANGLES = True
    yield
    E.ROPE2D_FP32_ANGLES = False


@pytest.fixture(scope="module")
def model():
    """Reduced-depth (1 PCB + 1 IRG, full 14B widths) fusion model on the GPU with the per-key synthetic weights."""
    import fwb200
    from fwb200.synth import build_fusion_model
    fwb200.require_device()
    m = build_fusion_model(num_dit_layers=2, start_index=1, device="cuda", seed=0, heads=True, gen_device="cpu")
    g = gold("joint_forward.pt")
    m.vggt.depth_head.intermediate_layer_idx = g["head_layer_idx"]
    m.vggt.point_head.intermediate_layer_idx = g["head_layer_idx"]
    return m


def _inputs(g, device="cuda", dtype=torch.bfloat16):
    from fwb200.synth import synth_inputs
    f, h, w = g["grid"]
    return synth_inputs(f, h, w, device=device, seed=1024, text_len=g["text_len"], dtype=dtype)


def test_state_dict_roundtrip_with_reference_schema(model):
    from _common import schema
    sd = model.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == schema()
    ref_sd = synth_state_dict()
    k = "IRGBlock.0.bicross_attention.cross_attn.m1_proj.weight"
    assert torch.equal(sd[k].float().cpu(), ref_sd[k].to(torch.bfloat16).float())


def test_irg_block_config1_vs_golden_and_oracle(model):
    """BASELINE config 1 on the GPU: single IRG block forward, f,h,w = 1,4,4."""
    import fwb200.engine as E
    from oracle import fw_oracle as O
    g = gold("irg_block_c1.pt")
    gen = torch.Generator().manual_seed(g["seed"])
    f, h, w = 1, 4, 4
    L = f * h * w
    x_dit = torch.randn(1, L, 5120, generator=gen)
    x_agg = torch.randn(f, 5 + h * w, 1024, generator=gen)
    context = torch.randn(1, 257 + g["text_len"], 5120, generator=gen)
    t_mod = torch.randn(1, 6, 5120, generator=gen) * 0.1
    e0 = torch.randn(1, 6, 1024, generator=gen) * 0.1
    plucker = torch.randn(1, L, 2048, generator=gen)
    dev = "cuda"
    fr, fd, fa = model.rope_tables(f, h, w, dev)
    pos = model.vggt.aggregator._positions(f, h, w, torch.device(dev))
    bf = torch.bfloat16

    def run_gpu(uncond=False):
        with torch.no_grad():
            return model.IRGBlock[0](x_dit=x_dit.to(dev, bf), x_agg=x_agg.to(dev, bf), context=context.to(dev, bf),
                                     t_mod=t_mod.to(dev, bf), freqs=fr, freqs_dit=fd, freqs_agg=fa, pos=pos, e0=e0.to(dev),
                                     uncond=uncond, plucker_fea=plucker.to(dev, bf), plucker_context_lens=torch.ones(1, dtype=torch.long))

    # (ii) against the reference's fp32 (no-autocast) output, with the same fp32 RoPE angles: bf16-sized error
    E.ROPE2D_FP32_ANGLES = True
    try:
        xd, xa, inter = run_gpu()
    finally:
        E.ROPE2D_FP32_ANGLES = False
    assert xd.dtype == torch.bfloat16 and xa.dtype == torch.float32   # the geometry stream is fp32 after modulation (Appendix A.3)
    assert inter[0].shape == (1, f, 5 + h * w, 1024)
    assert rel_err(xd.cpu(), g["x_dit_out"]) < 2e-2, rel_err(xd.cpu(), g["x_dit_out"])
    assert rel_err(xa.cpu(), g["x_agg_out"]) < 2e-2, rel_err(xa.cpu(), g["x_agg_out"])
    # (i) default mode against the oracle emulating the reference's CUDA-autocast rounding points (incl. bf16 RoPE angles)
    xd2, xa2, _ = run_gpu()
    sd = synth_state_dict()
    _, opos = O.aggregator_input(sd, "vggt.aggregator", torch.zeros(1, f, h, w, 1024))
    r = O.BF16.r
    od, oa, _ = O.irg_block(sd, "IRGBlock.0", r(x_dit), r(x_agg), r(context), r(t_mod), O.rope_table_3d(128, f, h, w),
                            O.rope_table_3d(96, f, h, w), O.rope_table_3d_with_extra(96, f, h, w, 5), opos, e0, r(plucker), nm=O.BF16)
    assert rel_err(xd2.cpu(), od) < 1e-2, rel_err(xd2.cpu(), od)
    assert rel_err(xa2.cpu(), oa) < 1e-2, rel_err(xa2.cpu(), oa)
    # never (much) less accurate than the emulated-bf16 reference itself (fp32-angle run vs fp32-angle emulation)
    assert rel_err(xd.cpu(), g["x_dit_out"]) < 1.5 * rel_err(od, g["x_dit_out"]) + 5e-3
    # uncond=True: the adapter is skipped (fusion/layer/block.py:70-72); own golden from the reference
    gu = gold("irg_block_c1_uncond.pt")
    E.ROPE2D_FP32_ANGLES = True
    try:
        xdu, xau, _ = run_gpu(uncond=True)
    finally:
        E.ROPE2D_FP32_ANGLES = False
    assert rel_err(xdu.cpu(), gu["x_dit_out"]) < 2e-2, rel_err(xdu.cpu(), gu["x_dit_out"])
    assert rel_err(xau.cpu(), gu["x_agg_out"]) < 2e-2, rel_err(xau.cpu(), gu["x_agg_out"])


def test_joint_forward_with_heads_vs_golden(model, fp32_rope_angles):
    g = gold("joint_forward.pt")
    inp = _inputs(g)
    f, h, w = g["grid"]
    lens = torch.ones(f, dtype=torch.long, device="cuda")
    lens[1:] = 4
    ts = torch.tensor([g["timestep"]], device="cuda", dtype=torch.bfloat16)
    taps = {}
    hooks = [model.pipe.dit.blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_pcb", o.clone())),
             model.vggt.aggregator.frame_blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__("after_frame", o.clone())),
             model.IRGBlock[0].register_forward_hook(lambda m, i, o: taps.update(after_irg_x=o[0].clone(), after_irg_tokens=o[1].clone()))]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out, pred = model.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], clip_feature=inp["clip_feature"],
                                        y=inp["y"], use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"],
                                        plucker_context_lens=lens, return_prediction=True)
    for hk in hooks:
        hk.remove()
    assert out.shape == g["out"].shape == (1, 16, f, 2 * h, 2 * w)
    stage = {k: rel_err(taps[k].float().cpu().reshape(v.shape), v) for k, v in g["taps"].items()}
    e = rel_err(out.cpu(), g["out"])
    print("stage rel errs vs reference fp32:", stage, "out", e)
    assert all(v < 3e-2 for v in stage.values()), stage
    assert e < 3e-2, e
    for k, ref in g["pred"].items():
        assert pred[k].shape == ref.shape, k                  # 5 frames = 4*(2-1)+1, 64x64 maps: index layout identical
        assert torch.isfinite(pred[k].float()).all(), k
    assert rel_err(pred["depth"].cpu(), g["pred"]["depth"]) < 8e-2
    assert rel_err(pred["depth_conf"].cpu(), g["pred"]["depth_conf"]) < 8e-2
    assert rel_err(pred["world_points_conf"].cpu(), g["pred"]["world_points_conf"]) < 8e-2
    assert rel_err(pred["pose_enc"].cpu(), g["pred"]["pose_enc"]) < 8e-2
    assert rel_err(pred["world_points"].cpu(), g["pred"]["world_points"]) < 8e-2, rel_err(pred["world_points"].cpu(), g["pred"]["world_points"])
    assert set(pred) == set(g["pred"])                         # every head output of the reference is produced and compared
    # second call reuses every hoisted invariant (context K/V, tables): identical result
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out2, none = model.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], clip_feature=inp["clip_feature"],
                                         y=inp["y"], use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"],
                                         plucker_context_lens=lens)
    assert none is None and torch.equal(out, out2)


def test_joint_forward_vs_bf16_oracle(model):
    from oracle import fw_oracle as O
    g = gold("joint_forward.pt")
    inp = _inputs(g)
    cpu = {k: v.float().cpu() for k, v in inp.items()}
    sd = synth_state_dict()
    ref, _, _ = O.joint_forward(sd, cpu["latents"], torch.tensor([g["timestep"]]), cpu["context_pos"], cpu["clip_feature"], cpu["y"],
                                cpu["plucker_fea"], start_index=1, n_irg=1, nm=O.BF16)
    ts = torch.tensor([g["timestep"]], device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        out, _ = model.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], clip_feature=inp["clip_feature"],
                                     y=inp["y"], use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"])
    e = rel_err(out.cpu(), ref)
    assert e < 2e-2, e


def test_denoise_step_and_sampler_loop(model, fp32_rope_angles):
    g = gold("denoise_step.pt")
    gj = gold("joint_forward.pt")
    inp = _inputs(gj)
    f, h, w = gj["grid"]
    # 50-step schedule, run steps 0..3 through generate_video's own loop body by calling it with 4 steps is not the same
    # schedule; so drive one step exactly as generate_video does.
    import fwb200
    sched = model.pipe.scheduler
    sched.set_timesteps(50)
    step = g["step"]
    t = sched.timesteps[step].unsqueeze(0).to(dtype=torch.bfloat16, device="cuda")
    lat = inp["latents"].clone()
    with torch.no_grad():
        p, _ = model.joint_forward(lat, timestep=t, context=inp["context_pos"], clip_feature=inp["clip_feature"], y=inp["y"],
                                   use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"])
        n, _ = model.joint_forward(lat, timestep=t, context=inp["context_neg"], clip_feature=inp["clip_feature"], y=inp["y"],
                                   use_gradient_checkpointing=False, plucker_fea=inp["plucker_fea"])
        fwb200.cfg_euler_step_(lat, p.contiguous(), n.contiguous(), 5.0, sched.dsigma(sched.timesteps[step]))
    assert rel_err(p.cpu(), g["pred_pos"]) < 3e-2 and rel_err(n.cpu(), g["pred_neg"]) < 3e-2
    assert rel_err(lat.cpu(), g["latents_next"]) < 1e-2
    # the public sampler API end to end (2 steps, injected latents, camera features from the pose encoder), heads last
    fwb200.reset_launch_count()
    plucker = torch.randn(1, 4 * (f - 1) + 1, 16 * h, 16 * w, 6, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        final, pred = model.generate_video(context_pos=inp["context_pos"], context_neg=inp["context_neg"],
                                           clip_feature=inp["clip_feature"], y=inp["y"], height=16 * h, width=16 * w,
                                           num_frames=4 * (f - 1) + 1, num_inference_steps=2, cfg_scale=5.0, seed=None,
                                           latents=inp["latents"], plucker_embedding=plucker)
    assert final.shape == inp["latents"].shape and torch.isfinite(final.float()).all()
    assert pred is not None and pred["depth"].shape == (1, 4 * (f - 1) + 1, 16 * h, 16 * w, 1)
    assert fwb200.launch_count() > 100   # our kernels did the work


def test_standalone_geometry_branch(model):
    """BASELINE config 5 analogue: VGGT.forward (aggregator without the adapter + heads) runs and is finite."""
    torch.manual_seed(0)
    # needs global blocks; the fusion surgery moved block 0 into the IRG block, so borrow it back for this check
    agg = model.vggt.aggregator
    saved, saved_n = agg.global_blocks[0], agg.aa_block_num
    agg.global_blocks[0] = model.IRGBlock[0].x_agg
    agg.aa_block_num = 1
    try:
        patch = torch.randn(1, 5120, 2, 4, 4, device="cuda", dtype=torch.bfloat16)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            pred = model.vggt(patch, t=torch.tensor([500.0], device="cuda"))
        assert pred["depth"].shape == (1, 5, 64, 64, 1) and pred["world_points"].shape == (1, 5, 64, 64, 3)
        assert pred["pose_enc"].shape == (1, 5, 9)
        assert all(torch.isfinite(v.float()).all() for v in pred.values())
    finally:
        agg.global_blocks[0] = saved
        agg.aa_block_num = saved_n


def test_pose_encoder_vs_reference_golden():
    """CameraPoseEncoder (once per sample): Plücker rays [1,5,64,64,6] -> camera features [1, 32, 2048]."""
    from FantasyWorld.diffsynth_wan21.models.pose_adaptor_ac3d import CameraPoseEncoder
    from fwb200.synth import synth_init
    g = gold("pose_encoder.pt")
    enc = CameraPoseEncoder(context_dim=2048, in_channels=6, downscale_coef=8, pose_inject_method="adaln")
    wrap = torch.nn.Module()
    wrap.camera_condition = torch.nn.Module()
    wrap.camera_condition.pose_encoder = enc
    assert {k: list(v.shape) for k, v in wrap.state_dict().items()} == g["schema"]
    synth_init(wrap, seed=0, gen_device="cpu")
    enc = enc.to("cuda").to(torch.bfloat16).eval()
    x = torch.randn(1, 5, 64, 64, 6, generator=torch.Generator().manual_seed(g["seed"]))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y = enc(x.to("cuda", torch.bfloat16))
    assert y.shape == g["out"].shape
    assert rel_err(y.cpu(), g["out"]) < 3e-2, rel_err(y.cpu(), g["out"])


def test_wan22_joint_forward_vs_reference_golden(fp32_rope_angles):
    """Wan2.2-Fun-A14B-Control-Camera fusion model (BASELINE config 4 family), reduced depth, GPU vs the reference golden."""
    import json
    import torch.nn as nn
    from _common import GOLD
    from FantasyWorld.diffsynth_wan22.models.wan_video_dit import WAN22_FUN_A14B_CONTROL_CAMERA
    from FantasyWorld.fusion.model_wan22 import FantasyWorldFusionModel as Fusion22
    from FantasyWorld.diffsynth_wan21.models.wan_video_dit import precompute_freqs_cis_3d
    from fwb200.synth import VGGT_CFG, materialize, synth_init, synth_inputs
    g = gold("joint_forward_wan22.pt")
    schema = json.loads((GOLD / "schema_wan22_reduced.json").read_text())
    with torch.device("meta"):
        m = Fusion22(start_index=1, use_gradient_checkpointing=False, cross_attention_list=[0], dit_path=None, lora_path=None,
                     vggt_cfg=dict(VGGT_CFG, enable_camera=False, enable_depth=False, enable_point=False), camera_control=True,
                     camera_cfg=dict(use_info="plucker"), dit_config=dict(WAN22_FUN_A14B_CONTROL_CAMERA, num_layers=2))
    agg = m.vggt.aggregator
    agg.frame_blocks = nn.ModuleList(list(agg.frame_blocks)[:1])
    agg.global_blocks = nn.ModuleList(list(agg.global_blocks)[:1])
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == schema      # same keys as the reference's Wan2.2 model
    materialize(m, "cuda", torch.bfloat16)
    m.pipe.dit.freqs = precompute_freqs_cis_3d(128)
    m.freqs_bicross = precompute_freqs_cis_3d(96)
    synth_init(m, seed=0, gen_device="cpu")
    m.eval()
    f, h, w = g["grid"]
    inp = synth_inputs(f, h, w, device="cuda", seed=1024, text_len=g["text_len"])
    control = torch.randn(1, 24, f, 16 * h, 16 * w, generator=torch.Generator().manual_seed(g["control_seed"])).to("cuda", torch.bfloat16)
    ts = torch.tensor([g["timestep"]], device="cuda", dtype=torch.bfloat16)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out, none = m.joint_forward(inp["latents"], timestep=ts, context=inp["context_pos"], y=inp["y"], use_gradient_checkpointing=False,
                                    control_camera_latents_input=control)
    assert none is None and out.shape == g["out"].shape
    e = rel_err(out.cpu(), g["out"])
    assert e < 3e-2, e


def test_geometry_heads_index_exact_on_cuda():
    """SURVEY §8 a19 on the CUDA path: which tokens reach which head stage (DPT layer selection, `[:, f0:f1, 5:]` slices,
    4-/16-frame chunking; camera token slice + 4x temporal expansion) — torch.equal against the reference's own record."""
    from _head_index import record_head_indexing
    from FantasyWorld.vggt.heads.camera_head import CameraHead
    from FantasyWorld.vggt.heads.dpt_head import DPTHead_3D_Causal
    from fwb_synth import synth_init
    g = gold("head_index.pt")
    wrap = torch.nn.Module()
    wrap.vggt = torch.nn.Module()
    wrap.vggt.depth_head = DPTHead_3D_Causal(dim_in=2048, output_dim=2, activation="exp", conf_activation="expp1", patch_size=16)
    wrap.vggt.camera_head = CameraHead(dim_in=2048)
    synth_init(wrap, seed=0, gen_device="cpu")
    wrap = wrap.to("cuda").to(torch.bfloat16).eval()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        rec = record_head_indexing(wrap.vggt.depth_head, wrap.vggt.camera_head, device="cuda", dtype=torch.bfloat16)
    gold_rec = g["records"]
    assert [(n, tuple(s)) for n, s, _ in rec] == [(n, tuple(s)) for n, s, _ in gold_rec]
    for (n, _, v), (_, _, gv) in zip(rec, gold_rec):
        if gv is not None:
            assert torch.equal(v, gv), n


def test_tiled_vae_decode_at_c2_size_on_cuda():
    """SURVEY §8f N1 on the GPU: the call inference_wan21.py:324-330 makes after the sampler — 21 latent frames of 60x104, tile
    (30, 52), stride (15, 26) -> 81 frames of 480x832 in [-1, 1]; fp32 vs the CPU golden-pinned math on a small case, and the
    bf16 full-size run finite with the right shape."""
    import time
    from _common import gold
    from FantasyWorld.diffsynth_wan21.pipelines.wan_video import WanVideoPipeline
    from fwb_synth import synth_init
    g = gold("vae.pt")
    pipe = WanVideoPipeline(device="cuda", torch_dtype=torch.bfloat16)
    vae = pipe.enable_vae(dtype=torch.float32)
    wrap = torch.nn.Module()
    wrap.vae = vae
    vae.model.requires_grad_(True)
    synth_init(wrap, seed=0, gen_device="cpu")
    vae.model.requires_grad_(False)
    z = torch.randn(1, 16, 3, 6, 8, generator=torch.Generator().manual_seed(11))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            tiled = vae.decode(z.cuda(), device="cuda", tiled=True, tile_size=g["tile_size"], tile_stride=g["tile_stride"])
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    assert rel_err(tiled.cpu(), g["tiled"]) < 1e-4, rel_err(tiled.cpu(), g["tiled"])
    vae.to(torch.bfloat16)
    lat = torch.randn(1, 16, 21, 60, 104, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        video = vae.decode(lat, device="cuda", tiled=True, tile_size=(30, 52), tile_stride=(15, 26))
    torch.cuda.synchronize()
    print(f"tiled VAE decode 81x480x832 on B200: {time.perf_counter() - t0:.2f} s (9 tiles, whole-clip convolutions)")
    assert video.shape == (1, 3, 81, 480, 832) and torch.isfinite(video.float()).all()
    assert float(video.max()) <= 1.0 and float(video.min()) >= -1.0
