"""Drop-in for the reference's validation module (training.py:19, HLVAE_main.py:21 import `validate`):
the reference file is loaded unchanged and its module-level `validation_dubo` (validation.py:16-76, which
calls the removed `torch.solve`) is rebound to hlvae_b200.validation.validation_dubo, so `validate`
(:78-260) picks it up through its own globals."""
import importlib.util
import os
import sys

from hlvae_b200.validation import validation_dubo  # noqa: F401

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _f = os.path.join(_p or ".", "validation.py")
    if os.path.exists(_f) and os.path.dirname(os.path.abspath(_f)) != _here:
        _spec = importlib.util.spec_from_file_location("_hlvae_reference_validation", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        _ref.validation_dubo = validation_dubo
        for _k, _v in vars(_ref).items():
            if not _k.startswith("__"):
                globals().setdefault(_k, _v)
        break
