"""skysampler-b200: the Gaussian-KDE score / draw hot path of vargatn/skysampler on NVIDIA B200.

Public surface (mirrors ``skysampler/emulator.py``):

    from skysampler_b200 import KDEContainer, GaussianKDE
    from skysampler_b200 import emulator      # make_*_infodicts, calc_scores*, run_scores*, accept_concentric, KFoldCV
    from skysampler_b200 import concentric    # batched concentric-shell driver;  pipeline: RatioSampler;  dist: multi-GPU
    from skysampler_b200 import features      # construct_wide/deep_container, FeatureSpaceContainer, DeepFeatureContainer

The compute lives in ``csrc/libskb.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/skb.h``).  There is no CPU fallback.
"""
from .kde import KDEContainer, GaussianKDE, weighted_mean, weighted_std  # noqa: F401
from . import emulator  # noqa: F401

__version__ = "0.1.0"
