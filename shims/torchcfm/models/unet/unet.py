"""`torchcfm.models.unet.unet.UNetModel` is the RAW guided-diffusion constructor (`image_size, in_channels,
model_channels, ...`: src/models/components/unet_4to3.py:5,51-67; configs/model/conditional_flow_matching_masked_condition.yaml:19)
-> stain2stain_b200.unet.RawUNetModel."""
from stain2stain_b200.unet import RawUNetModel as UNetModel  # noqa: F401

__all__ = ["UNetModel"]
