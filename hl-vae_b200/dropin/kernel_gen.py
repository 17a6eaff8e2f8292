"""Drop-in for the reference's kernel_gen module: generate_kernel_batched (kernel_gen.py:199-310, imported by
HLVAE_main.py:15) and the un-batched builders generate_kernel / generate_kernel_approx (:9-197), all assembled from
this repo's kernel classes (same module tree and parameter names as the gpytorch objects the reference builds)."""
from hlvae_b200.kernels import generate_kernel, generate_kernel_approx, generate_kernel_batched  # noqa: F401
