"""`from torchdyn.core import NeuralODE` (reference: src/models/conditional_flow_matching.py:7, configs/model/*.yaml
`solver._target_`) -> stain2stain_b200.neural_ode.NeuralODE."""
from stain2stain_b200.neural_ode import NeuralODE  # noqa: F401

__all__ = ["NeuralODE"]
