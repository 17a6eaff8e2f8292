"""Drop-in for HL_VAE/loglik.py: HLVAE.py:388 looks the functions up by name."""
from hlvae_b200.loglik import loglik_cat, loglik_count, loglik_ordinal, loglik_pos, loglik_real  # noqa: F401
