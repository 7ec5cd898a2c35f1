"""B200-native (sm_100a) implementation of the Xception -> LSTM -> head hot path of
Tonmoy1321/Multimodal-DeepFake-Detection.  See DESIGN.md and include/xcp.h.

Importing this package does not touch CUDA and does not load the shared library (DataLoader workers import it);
the first kernel call dlopen()s libxcp_sm100.so and fails loudly if it is missing."""
from .modules import (ArcFaceHead, AUFaceCrossDetector, BCELoss, Block, CBFocalLoss, FusedLinear, FusedLSTM, FusionHead,  # noqa: F401
                      LabelSmoothingBCEWithLogitsLoss,
                      SeparableConv2d, Xception, XceptionLSTMA, XceptionLSTMV, model_urls, xception)
from ._lib import XcpError, LIB_PATH  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .audio_frontend import MFCC  # noqa: F401
from .graph import GraphedInference, GraphedTrainStep, HostPrefetcher  # noqa: F401

__version__ = "0.1.0"
