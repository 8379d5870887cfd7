"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every symbol that
include/fwb200.h declares; the product path refuses to run without its kernels (no fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_loads_and_exports_all_declared_symbols():
    import fwb200
    names = fwb200.abi_symbols()
    assert len(names) >= 9 and "fwb_gemm_bf16" in names and "fwb_attn_fwd" in names
    lib = ctypes.CDLL(str(fwb200.library_path()))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fwb200.h but not exported"
    assert fwb200.lib.fwb_abi_version() == 2


def test_header_has_no_torch_types():
    text = (ROOT / "include" / "fwb200.h").read_text()
    assert "torch" not in text.lower().replace("pytorch", "") or "at::" not in text
    assert "at::Tensor" not in text and "c10::" not in text
    assert re.search(r'extern "C"', text)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_ops_fail_loudly_without_gpu():
    import fwb200
    assert not fwb200.device_ok()
    with pytest.raises(RuntimeError):
        fwb200.require_device()
    with pytest.raises(RuntimeError):
        fwb200.linear(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        fwb200.ln_modulate(torch.zeros(8, 8), eps=1e-6)


def test_argument_validation_messages():
    import fwb200
    from fwb200._lib import Epilogue, lib
    ep = Epilogue()
    rc = lib.fwb_gemm_bf16(None, 8, None, 8, 1, 8, 8, ctypes.byref(ep), None)
    assert rc != 0 and b"null" in lib.fwb_last_error()
    rc = lib.fwb_ln_modulate(ctypes.c_void_p(16), 0, 8, 4, 12, 1e-6, None, None, None, None, ctypes.c_void_p(16), 8, None)
    assert rc != 0 and b"multiple of 8" in lib.fwb_last_error()


def test_mirror_state_dict_schema_matches_reference():
    """state_dict keys are part of the ABI (the released .pth must load with no unexpected keys, inference_wan21.py:215-220).
    The schema fixture was written from the reference's own state_dict by tools/make_golden.py."""
    import json
    from fwb200.synth import CAMERA_CFG, VGGT_CFG, WAN21_I2V_14B
    from FantasyWorld.fusion.model_wan21 import FantasyWorldFusionModel
    import torch.nn as nn
    schema = json.loads((ROOT / "tests" / "golden" / "schema_reduced.json").read_text())
    with torch.device("meta"):
        m = FantasyWorldFusionModel(start_index=1, use_gradient_checkpointing=False, cross_attention_list=[0], dit_path=None,
                                    vggt_cfg=dict(VGGT_CFG), camera_control=True, camera_cfg=dict(CAMERA_CFG),
                                    dit_config=dict(WAN21_I2V_14B, num_layers=2))
    agg = m.vggt.aggregator
    agg.frame_blocks = nn.ModuleList(list(agg.frame_blocks)[:1])
    agg.global_blocks = nn.ModuleList(list(agg.global_blocks)[:1])
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == schema
    assert isinstance(m.pipe.dit.blocks[1], nn.Identity) and isinstance(agg.global_blocks[0], nn.Identity)
