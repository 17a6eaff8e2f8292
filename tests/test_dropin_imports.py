"""CPU: the drop-in wiring of INTEGRATION.md section 4.  With `hl-vae_b200/dropin` ahead of the reference
checkout on sys.path, the reference's own unmodified `training.py` / `HLVAE.py` must import and their
hot-path names must resolve to this repo's functions.  Needs /root/reference (build container only; the
GPU box skips it) and uses the test-only stand-ins for gpytorch / matplotlib, which are absent here."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HLVAE_REFERENCE", "/root/reference")

SCRIPT = r'''
import sys
root, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, root + "/hl-vae_b200/dropin", root + "/oracle/standins", ref]
import hlvae_b200
import training, HLVAE, validation, kernel_gen, kernel_spec, elbo_functions
from HL_VAE import loglik, read_functions
from hlvae_b200 import elbo as E, kernels as K, loglik as LL
assert training.__file__.startswith(ref)                                             # the reference's own file
from hlvae_b200 import theta as TH
assert "dropin" in HLVAE.__file__ and HLVAE.HLVAE.__module__ == "_hlvae_reference_HLVAE"   # reference class, passed through
assert HLVAE.HLVAE.theta_estimation is TH.theta_estimation and HLVAE.Observation_Cat.__module__ == "_hlvae_reference_HLVAE"
from hlvae_b200 import normalize as NZ
assert HLVAE.HLVAE.forward.__globals__["batch_normalization"] is NZ.batch_normalization
assert HLVAE.HLVAE.encode.__module__ == "_hlvae_reference_HLVAE" and callable(HLVAE.HLVAE.reference_theta_estimation)
assert training.minibatch_KLD_upper_bound is E.minibatch_KLD_upper_bound
assert training.minibatch_KLD_upper_bound_iter is E.minibatch_KLD_upper_bound_iter
assert kernel_gen.generate_kernel_batched is K.generate_kernel_batched
assert kernel_spec.CatKernel is K.CatKernel and kernel_spec.BinKernel is K.BinKernel
assert HLVAE.loglik is loglik and loglik.loglik_cat is LL.loglik_cat and loglik.loglik_ordinal is LL.loglik_ordinal
from hlvae_b200 import validation as VA0
assert validation.deviance_upper_bound is VA0.deviance_upper_bound and elbo_functions.elbo is VA0.elbo
assert kernel_gen.generate_kernel_approx is K.generate_kernel_approx and kernel_gen.generate_kernel is K.generate_kernel
assert training.read_functions is read_functions and hasattr(read_functions, "read_data")
assert read_functions.statistics.__module__.endswith("read_functions") and "hl-vae_b200" in read_functions.__file__
import utils as RU, model_test
from hlvae_b200 import predict as PR
assert RU.batch_predict_varying_T is PR.batch_predict_varying_T and model_test.batch_predict is PR.batch_predict
assert RU.HensmanDataLoader.__module__ == "_hlvae_reference_utils" and training.SubjectSampler is RU.SubjectSampler
from hlvae_b200 import validation as VA
assert validation.validation_dubo is VA.validation_dubo and validation.validate.__globals__["validation_dubo"] is VA.validation_dubo
assert training.validate is validation.validate
import HL_VAE.utils as U
assert U.__file__.startswith(ref)                                                     # everything else: reference
print("dropin ok")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_training_imports_resolve_to_dropins():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "dropin ok" in r.stdout, r.stdout + r.stderr
