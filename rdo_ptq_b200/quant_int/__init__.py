"""Drop-in for `from quant_int import *` (light-uniform-PTQ/quant_int/__init__.py)."""
from .quantizer import StraightThrough, round_ste, ActQuantizer, UniformAffineQuantizer     # noqa: F401
from .quant_layer import QuantModule                                                         # noqa: F401
from .quant_model import QuantModel, QuantCodingModel                                        # noqa: F401
