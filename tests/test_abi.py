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
    assert fwb200.lib.fwb_abi_version() == 3


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


def test_attention_tile_schedule_planner():
    """fwb_attn_plan is pure host arithmetic (no device): the tail split must leave the 1-GPU BASELINE shapes alone, split the
    sequence-parallel shard shapes, and always fit the workspace."""
    import ctypes as C
    import fwb200
    lib = fwb200.lib
    ws = int(4 * 148 * 256 * 129 * 4)                     # fwb_attn_workspace_bytes() on a 148-SM part

    def plan(B, H, Lq, Lk, D, ws_bytes=ws, sms=148):
        n_full, S = C.c_int(), C.c_int()
        assert lib.fwb_attn_plan(B, H, Lq, Lk, D, ws_bytes, sms, C.byref(n_full), C.byref(S)) == 0
        n_tiles = -(-Lq // 256) * H * B
        tail = n_tiles - n_full.value
        assert 0 <= n_full.value <= n_tiles and n_full.value % sms == 0 or S.value == 1
        if S.value > 1:
            assert 2 <= S.value <= 16 and Lk // S.value >= 512
            assert tail * S.value * 256 * (D + 1) * 4 <= ws_bytes
        else:
            assert n_full.value == n_tiles
        return n_tiles, n_full.value, S.value

    # 1 GPU, C2: DiT self-attention, adapter (both directions), VGGT global / frame, text cross-attention: unsplit
    for shape in [(1, 40, 32760, 32760, 128), (1, 12, 32760, 32865, 96), (1, 12, 32865, 32760, 96), (1, 16, 32865, 32865, 64),
                  (21, 16, 1565, 1565, 64), (1, 40, 32760, 512, 128), (1, 40, 32760, 257, 128)]:
        assert plan(*shape)[2] == 1, shape
    # 8 ranks: one K|V slice of the DiT self-attention (640 tiles = 4 waves + 48), adapter, VGGT global
    assert plan(1, 40, 4095, 8190, 128) == (640, 592, 3)
    n, full, S = plan(1, 12, 4095, 32865, 96)
    assert (n, full) == (192, 148) and S >= 3
    n, full, S = plan(1, 12, 4695, 32760, 96)
    assert (n, full) == (228, 148) and S >= 3
    n, full, S = plan(1, 16, 4695, 32865, 64)
    assert (n, full) == (304, 296) and S == 16
    # fewer tiles than SMs and long keys: everything is split; short keys: never
    assert plan(1, 2, 300, 5000, 128)[1:] == (0, 9)
    assert plan(1, 2, 300, 600, 128)[2] == 1
    # no workspace -> no split; tile count a multiple of the SM count -> no split
    assert plan(1, 40, 4095, 8190, 128, ws_bytes=0)[2] == 1
    assert plan(1, 37, 1024, 8192, 128)[2] == 1           # 4 * 37 = 148 tiles


def test_product_never_imports_test_infrastructure():
    """Contract: oracle/ (the CPU restatement + the staged reference) and tests/ are checkers.  Nothing under fantasy-world_b200/ may import
    them, and outside tests/ only __graft_entry__.smoke()/build(), bench.py (its CPU legs) and the golden / debug tools under tools/ do."""
    import ast
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    banned = ("oracle", "tests", "_ops_torch_shim", "_common", "ref_shim", "make_ref", "ref_runner", "fw_oracle")

    def imported_roots(path):
        out = set()
        for node in ast.walk(ast.parse(path.read_text())):
            if isinstance(node, ast.Import):
                out |= {a.name.split(".")[0] for a in node.names}
            elif isinstance(node, ast.ImportFrom) and node.module and node.level == 0:
                out.add(node.module.split(".")[0])
                out |= {a.name for a in node.names if node.module.split(".")[0] in banned}
        return out

    offenders = []
    for f in (root / "fantasy-world_b200").rglob("*.py"):
        if "build" in f.parts[len(root.parts):][1:2]:
            continue
        hit = imported_roots(f) & set(banned)
        if hit:
            offenders.append((str(f.relative_to(root)), sorted(hit)))
    assert not offenders, offenders
    # the C sources do not reference the oracle either (no CPU fallback is linked in)
    for f in (root / "fantasy-world_b200" / "csrc").glob("*"):
        assert "oracle" not in f.read_text().lower(), f
