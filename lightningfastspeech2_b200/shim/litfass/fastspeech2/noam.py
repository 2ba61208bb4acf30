"""litfass.fastspeech2.noam -> lightningfastspeech2_b200.fastspeech2.noam (same names, same signatures)"""
from lightningfastspeech2_b200.fastspeech2.noam import *  # noqa: F401,F403
from lightningfastspeech2_b200.fastspeech2 import noam as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
