"""Dense stand-in for the handful of gpytorch symbols the HL-VAE reference touches.

TEST INFRASTRUCTURE ONLY (oracle cross-check).  gpytorch is not
installed in this image and cannot be fetched; this package restates, from gpytorch's
documented behaviour, just enough of it for the *unmodified* reference modules
(kernel_spec.py, kernel_gen.py, GP_def.py, elbo_functions.py, training.py) to import and
run on CPU so that `oracle/hlvae_oracle.py` can be validated against them.  It is put on
sys.path only by `oracle/make_goldens.py` and by CPU tests that run the reference; the
product package never imports it.
"""
from . import constraints, distributions, kernels, likelihoods, means, models  # noqa: F401
