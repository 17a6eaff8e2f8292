"""hlvae_b200: B200-native implementation of HL-VAE's per-step ELBO hot path
(additive-kernel evaluation, KL upper bound sufficient statistics, heterogeneous masked
log-likelihoods, observation heads, batch normalisation, GP prediction / validation bound) behind the reference's Python operator surface.  See DESIGN.md."""
__version__ = "0.1.0"

from . import config  # noqa: F401


def __getattr__(name):
    import importlib
    if name in ("kernels", "likelihoods", "elbo", "loglik", "theta", "normalize", "predict", "validation", "graph",
                "subjects", "data", "parallel", "synth", "_lib"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
