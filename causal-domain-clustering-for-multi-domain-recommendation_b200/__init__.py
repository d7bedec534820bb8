"""B200-native (sm_100a) engine for the multi-domain CTR hot path of CDC-MDR.

The package mirrors the reference's Python model API (constructors, forward, get_regularization_loss, state_dict keys)
and executes everything through the C-ABI of libcdcmdr.so (include/cdcmdr.h).  Importing it without the built shared
library raises: there is no CPU or PyTorch fallback."""
from . import _lib  # noqa: F401
from .layer import BaseModel, FeaturesEmbedding, FeaturesLinear, MultiLayerPerceptron  # noqa: F401
from .optim import Adam  # noqa: F401
from .ple import PLE, CGC  # noqa: F401
from .mmoe import MMoE  # noqa: F401
from .dcn import DCN, DCNv2, CrossNetwork, CrossNetV2, CrossNetMix  # noqa: F401
from .star import STAR, MDR_BatchNorm, DNN  # noqa: F401
from .autoint import AutoInt  # noqa: F401
from .cdc import CDC  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from .data import DeviceLoader  # noqa: F401
from . import parallel  # noqa: F401
from . import metrics  # noqa: F401

__all__ = ["PLE", "CGC", "MMoE", "DCN", "DCNv2", "STAR", "AutoInt", "CDC", "Adam", "BaseModel", "GraphedTrainStep", "DeviceLoader"]
