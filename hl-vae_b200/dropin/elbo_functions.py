"""Drop-in for the reference's elbo_functions module (training.py:15 imports these two names)."""
from hlvae_b200.elbo import minibatch_KLD_upper_bound, minibatch_KLD_upper_bound_iter  # noqa: F401
