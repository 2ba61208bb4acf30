"""B200-native (sm_100a) mel-generation path of LightningFastSpeech2.

``lightningfastspeech2_b200.fastspeech2`` mirrors ``litfass.fastspeech2`` (same module /
class / state_dict names); the arithmetic runs in the hand-written CUDA kernels of
``liblfs2.so`` (C ABI: include/lfs2.h).  There is no CPU or library fallback.
"""
__version__ = "0.1.0"
