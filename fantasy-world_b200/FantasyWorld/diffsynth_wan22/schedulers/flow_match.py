"""Wan2.2 tree location of the flow-matching scheduler (same algorithm as diffsynth_wan21/schedulers/flow_match.py)."""
from ...diffsynth_wan21.schedulers.flow_match import FlowMatchScheduler  # noqa: F401
