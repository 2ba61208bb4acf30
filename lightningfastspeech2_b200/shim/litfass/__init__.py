"""Import shim: put ``lightningfastspeech2_b200/shim`` FIRST on sys.path (or call
``lightningfastspeech2_b200.install_shim()``) and ``litfass.fastspeech2.{fastspeech2,model,loss,noam}`` resolve to the
B200-native implementation, so litfass/train.py:24, generate.py:16 and synthesis/generator.py:20 run unedited.

Every other ``litfass`` sub-package (dataset, synthesis, third_party, ...) keeps resolving to an installed reference:
its directory, when one is found further down sys.path, is appended to this package's search path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _entry in list(sys.path):
    _cand = os.path.join(_entry or ".", "litfass")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.abspath(_cand) not in __path__:
        __path__.append(os.path.abspath(_cand))
