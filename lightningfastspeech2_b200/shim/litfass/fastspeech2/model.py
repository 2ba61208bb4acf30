"""litfass.fastspeech2.model -> lightningfastspeech2_b200.fastspeech2.model (same names, same signatures)"""
from lightningfastspeech2_b200.fastspeech2.model import *  # noqa: F401,F403
from lightningfastspeech2_b200.fastspeech2 import model as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
