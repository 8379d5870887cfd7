"""The reference runner (oracle/ref_runner.py: the UNMODIFIED reference in its own process) against the oracle on the CPU.

This pins the plumbing the GPU parity tests rely on: the runner builds the reference with the per-key synthetic weights and
the seeded block inputs of fwb_synth, and must therefore reproduce what the oracle computes from the same state_dict and
inputs (fp32, tolerance = summation order).  Skipped when no reference is staged (oracle/make_ref.py)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from _common import rel_err, synth_state_dict

ROOT = Path(__file__).resolve().parent.parent


def _available():
    return Path("/root/reference/FantasyWorld").exists() or (ROOT / "oracle" / "_ref" / "FantasyWorld").exists()


@pytest.mark.skipif(not _available(), reason="reference not staged (python oracle/make_ref.py)")
def test_staged_reference_is_unmodified():
    sys.path.insert(0, str(ROOT))
    from oracle import make_ref
    if (ROOT / "oracle" / "_ref" / "MANIFEST.json").exists():
        assert make_ref.verify()


@pytest.mark.skipif(not _available(), reason="reference not staged (python oracle/make_ref.py)")
def test_runner_blocks_match_oracle_fp32(tmp_path):
    from fwb_synth import synth_block_inputs
    from oracle import fw_oracle as O
    f, h, w, text_len = 1, 4, 4, 64
    r = subprocess.run([sys.executable, str(ROOT / "oracle" / "ref_runner.py"), "blocks", "--device", "cpu", "--grid", str(f), str(h), str(w),
                        "--text-len", str(text_len), "--modes", "fp32", "--out", str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["modes"]["fp32"]["dit_attention_backend"] == "F.scaled_dot_product_attention"
    ref = torch.load(tmp_path / "blocks_fp32.pt")
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth_state_dict().items()}      # the fp32 mode runs on bf16-valued weights
    inp = synth_block_inputs(f, h, w, text_len)
    tab, tab_d, tab_a = O.rope_table_3d(128, f, h, w), O.rope_table_3d(96, f, h, w), O.rope_table_3d_with_extra(96, f, h, w, 5)
    _, pos = O.aggregator_input(sd, "vggt.aggregator", torch.zeros(1, f, h, w, 1024))
    pcb = O.dit_block(sd, "pipe.dit.blocks.0", inp["x_dit"], inp["context"], inp["t_mod"], tab, inp["plucker"])
    frame = O.vggt_block(sd, "vggt.aggregator.frame_blocks.0", inp["x_agg"], pos, inp["e0"])
    xd, xa, _ = O.irg_block(sd, "IRGBlock.0", inp["x_dit"], inp["x_agg"], inp["context"], inp["t_mod"], tab, tab_d, tab_a, pos, inp["e0"],
                            inp["plucker"])
    assert rel_err(ref["pcb"], pcb) < 2e-4, rel_err(ref["pcb"], pcb)
    assert rel_err(ref["frame"], frame) < 2e-4, rel_err(ref["frame"], frame)
    assert rel_err(ref["irg_x"], xd) < 2e-4, rel_err(ref["irg_x"], xd)
    assert rel_err(ref["irg_tokens"].reshape(xa.shape), xa) < 2e-4, rel_err(ref["irg_tokens"].reshape(xa.shape), xa)
