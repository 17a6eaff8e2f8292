"""Shadow of the reference's HL_VAE package: `loglik` and `read_functions` come from here, every
other submodule (utils, ...) from the reference checkout further down sys.path."""
import os
import sys

for _p in sys.path:
    _cand = os.path.join(_p or ".", "HL_VAE")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)) \
            and os.path.exists(os.path.join(_cand, "utils.py")):
        __path__.append(_cand)
        break
