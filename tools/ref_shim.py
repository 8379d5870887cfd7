"""Compatibility import: the shim lives in oracle/ref_shim.py (test infrastructure, next to the oracle)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.ref_shim import *  # noqa: F401,F403,E402
from oracle.ref_shim import build_reference_fusion, import_reference, install_stubs, randomize_zero_init  # noqa: F401,E402
