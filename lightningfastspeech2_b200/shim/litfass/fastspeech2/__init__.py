"""``litfass.fastspeech2``: the hot-path modules come from lightningfastspeech2_b200; the reference's other
modules of this sub-package (fastdiff_variances, log_gmm, ...) keep resolving to an installed reference and
import ``litfass.fastspeech2.model`` -- i.e. this repo's classes -- like they always did."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _entry in list(sys.path):
    _cand = os.path.join(_entry or ".", "litfass", "fastspeech2")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.abspath(_cand) not in __path__:
        __path__.append(os.path.abspath(_cand))
