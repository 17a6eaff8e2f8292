"""Drop-in for the reference's kernel_gen module (HLVAE_main.py:15)."""
from hlvae_b200.kernels import generate_kernel_batched  # noqa: F401
