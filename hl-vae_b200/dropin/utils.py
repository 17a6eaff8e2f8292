"""Drop-in for the reference's utils module: everything (samplers, HensmanDataLoader, plotting, `predict`)
is passed through from the reference file found further down sys.path; the two batched GP posterior-mean
predictors (utils.py:99-191, :193-271 - dead on current torch, they call the removed `torch.solve`) are
served by hlvae_b200.predict."""
import importlib.util
import os
import sys

from hlvae_b200.predict import batch_predict, batch_predict_varying_T  # noqa: F401

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _f = os.path.join(_p or ".", "utils.py")
    if os.path.exists(_f) and os.path.dirname(os.path.abspath(_f)) != _here:
        _spec = importlib.util.spec_from_file_location("_hlvae_reference_utils", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        for _k, _v in vars(_ref).items():
            if not _k.startswith("__") and _k not in ("batch_predict", "batch_predict_varying_T"):
                globals().setdefault(_k, _v)
        break
