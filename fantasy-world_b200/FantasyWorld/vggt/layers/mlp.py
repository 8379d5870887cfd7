"""Mirror of FantasyWorld/vggt/layers/mlp.py: fc1 -> GELU(erf) -> fc2 on the fwb200 GEMM (GELU fused in the epilogue)."""
from typing import Callable, Optional

from torch import Tensor, nn

from fwb200 import engine as E
from fwb200 import ops


class Mlp(nn.Module):
    def __init__(self, in_features: int, hidden_features: Optional[int] = None, out_features: Optional[int] = None,
                 act_layer: Callable[..., nn.Module] = nn.GELU, drop: float = 0.0, bias: bool = True) -> None:
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop = nn.Dropout(drop)

    def forward(self, x: Tensor) -> Tensor:
        """ref: vggt/layers/mlp.py:34-40 (dropout is identity at inference)."""
        assert isinstance(self.act, nn.GELU) and self.act.approximate == "none"
        shp = x.shape
        h = E.lin(E.as_bf16(x).reshape(-1, shp[-1]), self.fc1, act=ops.ACT_GELU_ERF,
                  round_flags=ops.ROUND_AFTER_BIAS | ops.ROUND_AFTER_ACT)
        return E.lin(h, self.fc2, round_flags=ops.ROUND_AFTER_BIAS).view(*shp[:-1], -1)
