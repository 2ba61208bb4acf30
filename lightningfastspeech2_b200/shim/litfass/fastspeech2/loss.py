"""litfass.fastspeech2.loss -> lightningfastspeech2_b200.fastspeech2.loss (same names, same signatures)"""
from lightningfastspeech2_b200.fastspeech2.loss import *  # noqa: F401,F403
from lightningfastspeech2_b200.fastspeech2 import loss as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
