"""The oracle (oracle/fw_oracle.py, CPU restatement) against the golden vectors produced by the UNMODIFIED reference
(tools/make_golden.py).  This is what pins the oracle: same synthetic weights, same seeded inputs, fp32."""
import torch

from _common import gold, max_err, rel_err, synth_state_dict
from oracle import fw_oracle as O

TOL = 2e-4  # fp32 vs fp32, different summation orders over K = 5120..13824


def test_scheduler_matches_reference():
    g = gold("scheduler.pt")
    sig, ts = O.flow_match_sigmas(50)
    assert torch.equal(sig, g["sigmas"]) and torch.equal(ts, g["timesteps"])
    assert abs(float(ts[0]) - 1000.0) < 1e-3 and abs(float(ts[1]) - 995.9349) < 1e-2


def test_irg_block_config1():
    """BASELINE config 1: single IRG block forward, 1-frame 8x8 latent (4x4 tokens), CPU."""
    g = gold("irg_block_c1.pt")
    sd = synth_state_dict()
    gen = torch.Generator().manual_seed(g["seed"])
    f, h, w = 1, 4, 4
    L = f * h * w
    x_dit = torch.randn(1, L, 5120, generator=gen)
    x_agg = torch.randn(f, 5 + h * w, 1024, generator=gen)
    context = torch.randn(1, 257 + g["text_len"], 5120, generator=gen)
    t_mod = torch.randn(1, 6, 5120, generator=gen) * 0.1
    e0 = torch.randn(1, 6, 1024, generator=gen) * 0.1
    plucker = torch.randn(1, L, 2048, generator=gen)
    tab = O.rope_table_3d(128, f, h, w)
    tab_d = O.rope_table_3d(96, f, h, w)
    tab_a = O.rope_table_3d_with_extra(96, f, h, w, 5)
    _, pos = O.aggregator_input(sd, "vggt.aggregator", torch.zeros(1, f, h, w, 1024))
    xd, xa, _ = O.irg_block(sd, "IRGBlock.0", x_dit, x_agg, context, t_mod, tab, tab_d, tab_a, pos, e0, plucker)
    assert rel_err(xd, g["x_dit_out"]) < TOL, rel_err(xd, g["x_dit_out"])
    assert rel_err(xa, g["x_agg_out"]) < TOL, rel_err(xa, g["x_agg_out"])
    # uncond=True (fusion/layer/block.py:70-72): the bidirectional adapter is skipped; same inputs, own reference golden
    gu = gold("irg_block_c1_uncond.pt")
    xd, xa, _ = O.irg_block(sd, "IRGBlock.0", x_dit, x_agg, context, t_mod, tab, tab_d, tab_a, pos, e0, plucker, uncond=True)
    assert rel_err(xd, gu["x_dit_out"]) < TOL, rel_err(xd, gu["x_dit_out"])
    assert rel_err(xa, gu["x_agg_out"]) < TOL, rel_err(xa, gu["x_agg_out"])
    assert rel_err(gu["x_dit_out"], g["x_dit_out"]) > 1e-2     # the two branches really differ with the synthetic gammas


def _inputs(g):
    from fwb_synth import synth_inputs
    f, h, w = g["grid"]
    return synth_inputs(f, h, w, device="cpu", seed=1024, text_len=g["text_len"], dtype=torch.float32)


def test_joint_forward_reduced():
    g = gold("joint_forward.pt")
    sd = synth_state_dict()
    inp = _inputs(g)
    out, inter, patch = O.joint_forward(sd, inp["latents"], torch.tensor([g["timestep"]]), inp["context_pos"], inp["clip_feature"],
                                        inp["y"], inp["plucker_fea"], start_index=1, n_irg=1, collect_intermediates=True)
    assert out.shape == g["out"].shape
    assert rel_err(out, g["out"]) < TOL, rel_err(out, g["out"])
    # intermediates: [B, S, P, 2C] = frame block output | IRG geometry output
    taps = g["taps"]
    B, S, P, C2 = inter[0].shape
    assert rel_err(inter[0][..., :1024].reshape(S, P, 1024), taps["after_frame"]) < TOL
    assert rel_err(inter[0][..., 1024:].reshape(1, S * P, 1024), taps["after_irg_tokens"]) < TOL


def test_index_paths_bit_exact():
    """Integer / index work must be bit-identical: positions, token assembly order, unpatchify."""
    sd = synth_state_dict()
    f, h, w = 2, 3, 5
    patch = torch.arange(f * h * w * 1024, dtype=torch.float32).view(1, f, h, w, 1024)
    tokens, pos = O.aggregator_input(sd, "vggt.aggregator", patch)
    assert pos.dtype == torch.int64 and pos.shape == (f, 5 + h * w, 2)
    assert torch.equal(pos[:, :5], torch.zeros(f, 5, 2, dtype=torch.int64))
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    assert torch.equal(pos[0, 5:, 0], ys.reshape(-1) + 1) and torch.equal(pos[0, 5:, 1], xs.reshape(-1) + 1)
    assert torch.equal(tokens[:, 5:], patch.view(f, h * w, 1024))
    assert torch.equal(tokens[0, 0], sd["vggt.aggregator.camera_token"][0, 0, 0]) and torch.equal(tokens[1, 0], sd["vggt.aggregator.camera_token"][0, 1, 0])
    x = torch.arange(1 * f * h * w * 64, dtype=torch.float32).view(1, f * h * w, 64)
    u = O.unpatchify(x, (f, h, w))
    assert u.shape == (1, 16, f, 2 * h, 2 * w)
    # element (token (ff,hh,ww), (y,z,c)) lands at [c, ff, 2hh+y, 2ww+z]
    ff, hh, ww, yy, zz, cc = 1, 2, 3, 1, 0, 7
    assert u[0, cc, ff, 2 * hh + yy, 2 * ww + zz] == x[0, (ff * h + hh) * w + ww, (yy * 2 + zz) * 16 + cc]


def test_denoise_step():
    g = gold("denoise_step.pt")
    gj = gold("joint_forward.pt")
    sd = synth_state_dict()
    inp = _inputs(gj)
    sig, ts = O.flow_match_sigmas(50)
    nxt = O.denoise_step(sd, inp["latents"], g["step"], sig, ts.to(torch.bfloat16).float(), inp["context_pos"], inp["context_neg"],
                         inp["clip_feature"], inp["y"], inp["plucker_fea"], start_index=1, n_irg=1)
    assert rel_err(nxt, g["latents_next"]) < TOL, rel_err(nxt, g["latents_next"])


def test_bf16_emulation_stays_close_to_fp32():
    """The bf16-rounding emulation (what the CUDA path is compared with) must stay within bf16 noise of the fp32 oracle."""
    gj = gold("joint_forward.pt")
    sd = synth_state_dict()
    inp = _inputs(gj)
    out, _, _ = O.joint_forward(sd, inp["latents"], torch.tensor([gj["timestep"]]), inp["context_pos"], inp["clip_feature"],
                                inp["y"], inp["plucker_fea"], start_index=1, n_irg=1, nm=O.BF16)
    assert rel_err(out, gj["out"]) < 3e-2


def test_wan22_joint_forward_reduced():
    """Wan2.2-Fun-A14B-Control-Camera variant (no CLIP, control adapter): oracle vs the reference's model_wan22.joint_forward."""
    import json
    from _common import GOLD
    from fwb_synth import synth_inputs, synth_tensor
    g = gold("joint_forward_wan22.pt")
    schema = json.loads((GOLD / "schema_wan22_reduced.json").read_text())
    sd = {k: synth_tensor(k, shape, 0, "cpu") for k, shape in schema.items()}
    f, h, w = g["grid"]
    inp = synth_inputs(f, h, w, device="cpu", seed=1024, text_len=g["text_len"], dtype=torch.float32)
    control = torch.randn(1, 24, f, 16 * h, 16 * w, generator=torch.Generator().manual_seed(g["control_seed"]))
    out, _, _ = O.joint_forward(sd, inp["latents"], torch.tensor([g["timestep"]]), inp["context_pos"], None, inp["y"], None,
                                start_index=1, n_irg=1, control=control)
    assert rel_err(out, g["out"]) < TOL, rel_err(out, g["out"])
