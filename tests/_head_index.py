"""Shared by tools/make_golden_head_index.py (reference side) and the tests (mirror side): integer-coded token lists and the
hook-based recorder of which tokens reach which head stage."""
import torch

S, GH, GW, N_LAYERS, PATCH_START = 6, 2, 3, 24, 5        # 6 latent frames -> chunks 4 + 2; 21 video frames -> chunks 16 + 5


def coded_tokens(device="cpu", dtype=torch.float32):
    """24 x [1, S, 5 + GH*GW, 2048]: token (layer l, frame s, position p) carries the integer code in channel 0 (and code + c % 5
    in channel c, so LayerNorm sees a non-constant vector).  All values are small integers: exact in bf16."""
    P = PATCH_START + GH * GW
    out = []
    for layer in range(N_LAYERS):
        s = torch.arange(S).view(1, S, 1, 1)
        p = torch.arange(P).view(1, 1, P, 1)
        c = torch.arange(2048).view(1, 1, 1, 2048)
        code = ((layer * 31 + s * 7 + p * 3) % 127 - 63) + (c % 5)
        out.append(code.to(device=device, dtype=dtype))
    return out


def record_head_indexing(depth_head, camera_head=None, device="cpu", dtype=torch.float32):
    """Run the two heads on the coded tokens and return the list of hook records [(name, shape, int16 tensor or None)]."""
    rec = []
    hooks = []

    def ints(x):
        return x[..., 0].detach().round().to(torch.int16).cpu()

    def pre(name, with_values):
        def fn(mod, args):
            x = args[0]
            rec.append((name, tuple(x.shape), ints(x) if with_values else None))
        return fn

    hooks.append(depth_head.norm.register_forward_pre_hook(pre("dpt.norm", True)))
    hooks.append(depth_head.scratch.layer1_rn.register_forward_pre_hook(pre("dpt.scratch.layer1_rn", False)))
    hooks.append(depth_head.scratch.refinenet4.register_forward_pre_hook(pre("dpt.scratch.refinenet4", False)))
    hooks.append(depth_head.scratch.output_conv2.register_forward_pre_hook(pre("dpt.scratch.output_conv2", False)))
    if camera_head is not None:
        hooks.append(camera_head.camera_time_upsample.register_forward_pre_hook(pre("cam.camera_time_upsample", True)))
        hooks.append(camera_head.token_norm.register_forward_pre_hook(pre("cam.token_norm", True)))
        hooks.append(camera_head.adaln_norm.register_forward_pre_hook(pre("cam.adaln_norm", False)))
    toks = coded_tokens(device, dtype)
    images = torch.zeros(1, S, GH, GW, 1024, device=device, dtype=dtype)
    with torch.no_grad():
        depth, conf = depth_head(toks, images=images, patch_start_idx=PATCH_START)
        poses = camera_head(toks) if camera_head is not None else None
    for h in hooks:
        h.remove()
    rec.append(("dpt.out", tuple(depth.shape), None))
    rec.append(("dpt.conf", tuple(conf.shape), None))
    if poses is not None:
        rec.append(("cam.out", tuple(poses[-1].shape), None))
    return rec


