from .mlp import Mlp
from .attention import Attention
from .block import Block, CamTokenProjector
from .layer_scale import LayerScale
