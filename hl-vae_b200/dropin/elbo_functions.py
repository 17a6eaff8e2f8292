"""Drop-in for the reference's elbo_functions module.

training.py:15 imports `minibatch_KLD_upper_bound` and `minibatch_KLD_upper_bound_iter` (the per-step hot path);
validation.py:10 imports `deviance_upper_bound` and `elbo` (the un-batched, full-data evaluation bounds,
elbo_functions.py:9-115).  All four are served by the CUDA kernels; nothing is passed through from the reference
(its `elbo` / `deviance_upper_bound` call `torch.solve`, which current torch no longer has)."""
from hlvae_b200.elbo import minibatch_KLD_upper_bound, minibatch_KLD_upper_bound_iter  # noqa: F401
from hlvae_b200.validation import deviance_upper_bound, elbo  # noqa: F401
