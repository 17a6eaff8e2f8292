"""Drop-in for the reference's HLVAE module (`from HLVAE import HLVAE`, HLVAE_main.py:25).

The reference file found further down sys.path is executed unchanged (its `from HL_VAE import loglik` already
resolves to the CUDA per-type likelihoods) and every name is passed through; the model class then gets
  * `HLVAE.theta_estimation` (HLVAE.py:416-453)  -> hlvae_b200.theta.theta_estimation (one head kernel per
    direction instead of two einsum passes per type group + boolean-index scatters; same values and gradients
    for 0/1 masks), and
  * the module-level `batch_normalization` the methods call (HLVAE.py:7,303,330,370; HL_VAE/utils.py:88-143)
    -> hlvae_b200.normalize.batch_normalization (three streaming launches), and
  * `HLVAE.loglik_and_reconstruction` (HLVAE.py:381-414) -> hlvae_b200.loglik.loglik_and_reconstruction (all type
    groups in one launch) when `hlvae_b200.config.fused_loglik_method` is set - off by default because the fused
    method returns no `samples` (training.py never reads them; predict / test scripts may).
The encoder / decoder trunk stays the reference's stock PyTorch code."""
import importlib.util
import os
import sys

from hlvae_b200 import config as _config
from hlvae_b200.loglik import loglik_and_reconstruction as _fused_loglik_method
from hlvae_b200.normalize import batch_normalization as _batch_normalization
from hlvae_b200.theta import theta_estimation as _theta_estimation

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _f = os.path.join(_p or ".", "HLVAE.py")
    if os.path.exists(_f) and os.path.dirname(os.path.abspath(_f)) != _here:
        _spec = importlib.util.spec_from_file_location("_hlvae_reference_HLVAE", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        for _k, _v in vars(_ref).items():
            if not _k.startswith("__"):
                globals().setdefault(_k, _v)
        # `from HL_VAE.utils import batch_normalization` (HLVAE.py:7) -> the module global the methods call
        _ref.reference_batch_normalization = _ref.batch_normalization
        _ref.batch_normalization = batch_normalization = _batch_normalization
        _ref.HLVAE.reference_theta_estimation = _ref.HLVAE.theta_estimation
        _ref.HLVAE.theta_estimation = _theta_estimation
        if getattr(_config, "fused_loglik_method", False):
            _ref.HLVAE.reference_loglik_and_reconstruction = _ref.HLVAE.loglik_and_reconstruction
            _ref.HLVAE.loglik_and_reconstruction = _fused_loglik_method
        break
else:
    raise ImportError("hlvae_b200 drop-in HLVAE: the reference's HLVAE.py was not found on sys.path")
