from .layers import *            # noqa: F401,F403
from .layers import f_gdn        # noqa: F401
from .entropy_models import EntropyBottleneck, GaussianConditional     # noqa: F401
from .models import ScaleHyperprior, MeanScaleHyperprior, Cheng2020Attention, ARCHS   # noqa: F401
from .swin import WindowAttention, SwinTransformerBlock, BasicLayer, RSTB, PatchEmbed, PatchUnEmbed   # noqa: F401
from .nic import NIC   # noqa: F401
