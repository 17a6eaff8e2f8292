"""Drop-in for the reference's kernel_spec module (kernel_spec.py:9-69)."""
from hlvae_b200.kernels import BinKernel, CatKernel, RbfKernel, RBFKernel  # noqa: F401
