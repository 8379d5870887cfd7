"""Mirror of FantasyWorld/vggt/heads/head_act.py: output activations of the camera / depth / point heads."""
import torch
import torch.nn.functional as F


def inverse_log_transform(y):
    """sign(y) * (exp(|y|) - 1).  ref: head_act.py:114-125."""
    return torch.sign(y) * torch.expm1(torch.abs(y))


_POSE_ACTS = {"linear": lambda t: t, "inv_log": inverse_log_transform, "exp": torch.exp, "relu": F.relu}


def base_pose_act(pose_enc, act_type="linear"):
    if act_type not in _POSE_ACTS:
        raise ValueError(f"Unknown act_type: {act_type}")
    return _POSE_ACTS[act_type](pose_enc)


def activate_pose(pred_pose_enc, trans_act="linear", quat_act="linear", fl_act="linear"):
    """[..., 9] = translation(3) | quaternion(4) | field of view(2), each with its own activation.  ref: head_act.py:11-36."""
    return torch.cat([base_pose_act(pred_pose_enc[..., :3], trans_act), base_pose_act(pred_pose_enc[..., 3:7], quat_act),
                      base_pose_act(pred_pose_enc[..., 7:], fl_act)], dim=-1)


def activate_head(out, activation="norm_exp", conf_activation="expp1"):
    """out [B, C, H, W] -> (values [B, H, W, C-1], confidence [B, H, W]).  ref: head_act.py:61-112."""
    fmap = out.permute(0, 2, 3, 1)
    xyz, conf = fmap[..., :-1], fmap[..., -1]
    if activation == "norm_exp":
        d = xyz.norm(dim=-1, keepdim=True).clamp(min=1e-8)
        pts = xyz / d * torch.expm1(d)
    elif activation == "norm":
        pts = xyz / xyz.norm(dim=-1, keepdim=True)
    elif activation == "exp":
        pts = torch.exp(xyz)
    elif activation == "relu":
        pts = F.relu(xyz)
    elif activation == "inv_log":
        pts = inverse_log_transform(xyz)
    elif activation == "xy_inv_log":
        xy, z = xyz.split([2, 1], dim=-1)
        z = inverse_log_transform(z)
        pts = torch.cat([xy * z, z], dim=-1)
    elif activation == "sigmoid":
        pts = torch.sigmoid(xyz)
    elif activation == "linear":
        pts = xyz
    else:
        raise ValueError(f"Unknown activation: {activation}")
    if conf_activation == "expp1":
        c = 1 + conf.exp()
    elif conf_activation == "expp0":
        c = conf.exp()
    elif conf_activation == "sigmoid":
        c = torch.sigmoid(conf)
    else:
        raise ValueError(f"Unknown conf_activation: {conf_activation}")
    return pts, c
