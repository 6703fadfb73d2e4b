"""Drop-in for `from quantization import *` (task-oriented-PTQ/quantization/__init__.py:1-6)."""
from .quantizer import (StraightThrough, round_ste, lp_loss, ActQuantizer, ActQuant, UniformAffineQuantizer,   # noqa
                        AdaRoundQuantizer)
from .quant_layer import QuantModule, f_gdn                                                  # noqa: F401
from .quant_block import BaseQuantBlock, QuantRBWS, QuantRBU, QuantRB, QuantSC, QuantMlp, specials     # noqa: F401
from .quant_block import QuantWindowAttention, QuantSwinTransformerBlock, QuantBasicLayer, QuantRSTB   # noqa: F401
from .quant_model import QuantModel                                                          # noqa: F401
from .utils import LinearTempDecay, save_inp_oup_data, GetLayerInpOut, DataSaverHook, StopForwardException  # noqa
from .recon import DrawPlan, UnitTrainer, run_reconstruction                                 # noqa: F401
from .layer_opt import layer_reconstruction, find_unquantized_module                         # noqa: F401
from .block_opt import block_reconstruction                                                  # noqa: F401
