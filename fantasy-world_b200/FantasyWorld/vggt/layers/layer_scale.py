"""Mirror of FantasyWorld/vggt/layers/layer_scale.py (state-dict key: `gamma`).

Per-channel output scale of the VGGT residual branches.  On the fused block path gamma never runs as its own op: it is folded
into the epilogue of the proj / fc2 GEMM (`fwb_gemm_bf16` scale1 / scale2, see fwb200/engine.py); `forward` below is the
standalone form kept for API parity with the reference module."""
import torch
from torch import nn


class LayerScale(nn.Module):
    def __init__(self, dim, init_values=1e-5, inplace=False):
        super().__init__()
        start = torch.full((int(dim),), 1.0)
        self.gamma = nn.Parameter(start * init_values)       # init_values: float or a [dim] tensor
        self.inplace = bool(inplace)

    def forward(self, x):
        if self.inplace:
            x *= self.gamma
            return x
        return torch.mul(x, self.gamma)
