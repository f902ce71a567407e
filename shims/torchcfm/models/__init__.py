from . import unet  # noqa: F401
