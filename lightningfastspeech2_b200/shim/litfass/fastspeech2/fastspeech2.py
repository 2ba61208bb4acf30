"""litfass.fastspeech2.fastspeech2 -> lightningfastspeech2_b200.fastspeech2.fastspeech2 (same names, same signatures)"""
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import *  # noqa: F401,F403
from lightningfastspeech2_b200.fastspeech2 import fastspeech2 as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
