"""Run-time switches of the host layer."""

# Read the device status words (one small device->host copy) after the KL kernels and raise
# RuntimeError on a non positive-definite block, like torch.linalg.cholesky does in the reference
# path (elbo_functions.py:154-157).  The benchmark turns this off to keep the step sync-free.
check_errors = True

# torch.distributed process group over which the per-latent accumulators are all-reduced
# (None: single process).  Set by hlvae_b200.parallel.enable().
process_group = None

# Run independent kernels of one KL call on two CUDA streams (per-subject T x T stage next to the
# M x M pre-stage).  Pure scheduling: results are identical either way.
overlap = True

# Drop-in HLVAE module (hl-vae_b200/dropin/HLVAE.py): also replace HLVAE.loglik_and_reconstruction
# (HLVAE.py:381-414) by the fused method, which returns no `samples`.  Read when the drop-in is imported.
fused_loglik_method = False

# Keep the per-term pieces of the last KL call (the scalars A, B + sum(iB*K0), C, F per latent dimension, the
# M x M pre-stage scalars, S, p, iK, G) on `hlvae_b200.elbo.last_terms` - what tests compare term by term with
# elbo_functions.py:166-181 / :256-277.  Off by default (a few device copies per call).
keep_terms = False
