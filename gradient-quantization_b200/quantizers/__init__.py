"""Drop-in for the reference's `quantizers` package (quantizers/__init__.py:1)."""
from .base_quantizer import Quantizer
from .ps_quantizer import PSQuantizer
from .ring_quantizer import RingQuantizer

__all__ = ["Quantizer", "PSQuantizer", "RingQuantizer"]
