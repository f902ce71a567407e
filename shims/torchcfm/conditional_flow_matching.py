"""`from torchcfm.conditional_flow_matching import ConditionalFlowMatcher` (reference:
src/models/conditional_flow_matching.py:6; yaml `flow_matcher._target_`) -> stain2stain_b200.flow_matching."""
from stain2stain_b200.flow_matching import ConditionalFlowMatcher  # noqa: F401

__all__ = ["ConditionalFlowMatcher"]
