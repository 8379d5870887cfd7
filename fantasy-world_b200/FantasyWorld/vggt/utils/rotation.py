"""Quaternion <-> rotation-matrix helpers of the geometry branch's pose encoding (mirror of FantasyWorld/vggt/utils/rotation.py:
same names, same conventions — quaternions are (x, y, z, w), scalar LAST, real part made non-negative).  Host-side, once per
sample; plain torch.  Written as matrix identities rather than the reference's stacked-candidate formulation."""
from __future__ import annotations

import torch


def standardize_quaternion(quaternions: torch.Tensor) -> torch.Tensor:
    """Flip the sign so that the real part (last component) is >= 0."""
    return torch.where(quaternions[..., 3:4] < 0, -quaternions, quaternions)


def quat_to_mat(quaternions: torch.Tensor) -> torch.Tensor:
    """(..., 4) scalar-last quaternions (not necessarily unit) -> (..., 3, 3).  R = I + 2/|q|^2 * (w [v]x + [v]x^2)."""
    x, y, z, w = quaternions.unbind(-1)
    s = 2.0 / (quaternions * quaternions).sum(-1)
    rows = (1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
            s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
            s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y))
    return torch.stack(rows, dim=-1).reshape(quaternions.shape[:-1] + (3, 3))


def mat_to_quat(matrix: torch.Tensor) -> torch.Tensor:
    """(..., 3, 3) rotation matrices -> (..., 4) scalar-last quaternions with non-negative real part.

    Numerically robust branch selection: of the four quantities 4w^2, 4x^2, 4y^2, 4z^2 (each 1 +- traces) pick the largest as
    the pivot and derive the other three components from the off-diagonal sums / differences divided by 2 * sqrt(pivot)
    (pivot magnitude floored at 0.1 before the division, as the reference does)."""
    if matrix.shape[-2:] != (3, 3):
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    m = matrix
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    four_sq = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], dim=-1)   # w, x, y, z
    mag = torch.sqrt(four_sq.clamp_min(0))
    a, b, c = m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]     # 4wx, 4wy, 4wz
    d, e, f = m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0], m[..., 2, 1] + m[..., 1, 2]     # 4xy, 4xz, 4yz
    sq = mag * mag
    # candidate (w, x, y, z) * 2 * mag_pivot for each pivot
    cands = torch.stack([torch.stack([sq[..., 0], a, b, c], -1), torch.stack([a, sq[..., 1], d, e], -1),
                         torch.stack([b, d, sq[..., 2], f], -1), torch.stack([c, e, f, sq[..., 3]], -1)], dim=-2)
    cands = cands / (2.0 * mag.clamp_min(0.1))[..., None]
    pivot = mag.argmax(dim=-1)
    wxyz = torch.gather(cands, -2, pivot[..., None, None].expand(*pivot.shape, 1, 4)).squeeze(-2)
    return standardize_quaternion(wxyz[..., [1, 2, 3, 0]])
