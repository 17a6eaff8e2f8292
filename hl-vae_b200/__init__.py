"""hlvae_b200: B200-native implementation of HL-VAE's per-step ELBO hot path."""
__version__ = "0.1.0"
