"""`torchcfm.models.unet.UNetModel` is torchcfm's WRAPPER class (`dim=[C,H,W], num_channels, ...`:
configs/model/conditional_flow_matching.yaml:16-26) -> stain2stain_b200.unet.UNetModel."""
from stain2stain_b200.unet import UNetModel  # noqa: F401
from . import unet  # noqa: F401

UNetModelWrapper = UNetModel
__all__ = ["UNetModel", "UNetModelWrapper"]
