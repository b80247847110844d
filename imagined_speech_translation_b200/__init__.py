"""B200-native (sm_100a) hot path for imagined-speech EEG training.

Host-side mirror of the reference's data-transform / model-module interface
over the C ABI of libeegx.so (include/eegx.h).  No CPU fallback.
"""
from ._lib import EegxError, LIB_PATH  # noqa: F401
from .preprocess import (DSP_CONFIG, DSP_CONFIG_LONG, REGION_ORDER, RegionNormalizer,  # noqa: F401
                         SpectrogramFrontEnd, design_bandpass_fir, normalize_dense)

__version__ = "0.1.0"
from .layers import Conv1DWithAttention, FeedForwardNetwork, SqueezeExciteBlock  # noqa: F401,E402
from .brain_encoder import BrainRegionEncoder  # noqa: F401,E402
from .optim import FlatAdamW  # noqa: F401,E402
from .model import BARTDecoder, EEGDecodingModel  # noqa: F401,E402
from .data import EEGDataset, PrefetchLoader, TrialStore, apply_augmentation, augment_regions, build_region_indices  # noqa: F401,E402
from . import distributed, fused, ops  # noqa: F401,E402
