"""Drop-in for the reference's elbo_functions module.

training.py:15 imports `minibatch_KLD_upper_bound` and `minibatch_KLD_upper_bound_iter` (the hot
path, served by the CUDA kernels); validation.py:10 imports `deviance_upper_bound` and `elbo`
(full-data validation variants, out of scope), which are passed through from the reference module
found further down sys.path so that `import training` keeps working unchanged."""
import importlib.util
import os
import sys

from hlvae_b200.elbo import minibatch_KLD_upper_bound, minibatch_KLD_upper_bound_iter  # noqa: F401

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _f = os.path.join(_p or ".", "elbo_functions.py")
    if os.path.exists(_f) and os.path.dirname(os.path.abspath(_f)) != _here:
        _spec = importlib.util.spec_from_file_location("_hlvae_reference_elbo_functions", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        for _k, _v in vars(_ref).items():
            if not _k.startswith("_") and _k not in ("minibatch_KLD_upper_bound", "minibatch_KLD_upper_bound_iter"):
                globals().setdefault(_k, _v)
        break
