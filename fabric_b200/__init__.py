"""fabric_b200: the BiDateNet change-detection hot path of granularai/fabric, re-built for B200 (sm_100a).

Public surface (mirrors the reference's modules):
    fabric_b200.BiDateNet                     <- models/bidate_model.py
    fabric_b200.unet_parts.{double_conv,inconv,down,up,outconv}   <- models/unet_parts.py
The same classes are importable under the reference's own module paths ``models.bidate_model`` /
``models.unet_parts`` through the shim package ``models/`` at the repo root.
"""
from ._lib import FabricB200Error, LIB_PATH  # noqa: F401
from .bidate_model import BiDateNet  # noqa: F401
from . import unet_parts  # noqa: F401

__version__ = "0.1.0"
