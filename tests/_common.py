"""Shared helpers for the tests: synthetic state_dict from the committed schema, golden loading."""
from __future__ import annotations

import functools
import json
from pathlib import Path

import torch

GOLD = Path(__file__).resolve().parent / "golden"


def gold(name):
    return torch.load(GOLD / name, map_location="cpu", weights_only=False)


@functools.lru_cache(maxsize=1)
def schema():
    return json.loads((GOLD / "schema_reduced.json").read_text())


@functools.lru_cache(maxsize=1)
def synth_state_dict():
    """fp32 CPU state_dict with the reference's key schema (reduced depth: 1 PCB + 1 IRG), values from the per-key
    seeded generator — identical to what tools/make_golden.py loaded into the reference."""
    from fwb200.synth import synth_tensor
    return {k: synth_tensor(k, shape, seed=0, device="cpu") for k, shape in schema().items()}


def subset(sd, prefixes):
    return {k: v for k, v in sd.items() if any(k.startswith(p) for p in prefixes)}


def rel_err(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def max_err(a, b):
    return float((a.float() - b.float()).abs().max())
