"""Drop-in for reference models/bidate_model.py -- see fabric_b200/bidate_model.py."""
from fabric_b200.bidate_model import BiDateNet  # noqa: F401
from .unet_parts import down, outconv, up, inconv  # noqa: F401
