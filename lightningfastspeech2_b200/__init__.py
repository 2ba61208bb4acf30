"""B200-native (sm_100a) mel-generation path of LightningFastSpeech2.

``lightningfastspeech2_b200.fastspeech2`` mirrors ``litfass.fastspeech2`` (same module /
class / state_dict names); the arithmetic runs in the hand-written CUDA kernels of
``liblfs2.so`` (C ABI: include/lfs2.h).  There is no CPU or library fallback.
"""
__version__ = "0.1.0"


def install_shim():
    """Make ``import litfass.fastspeech2.fastspeech2`` (what litfass/train.py, generate.py and
    synthesis/generator.py import) resolve to this package: puts the bundled shim directory first on sys.path.
    Equivalent to PYTHONPATH=<repo>/lightningfastspeech2_b200/shim:<repo>."""
    import os
    import sys

    shim = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    for name in [n for n in sys.modules if n == "litfass" or n.startswith("litfass.fastspeech2")]:
        del sys.modules[name]
    return shim
