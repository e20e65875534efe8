"""Drop-in for reference models/unet_parts.py -- see fabric_b200/unet_parts.py."""
from fabric_b200.unet_parts import double_conv, inconv, down, up, outconv  # noqa: F401
