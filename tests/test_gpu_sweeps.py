"""Short, fixed-seed runs of the randomised parity sweeps under profiles/tools/ (the long runs are logged in
profiles/r02zb_*_stress.log): every case compares the CUDA path with the float64 oracle on a random configuration -
KL bound (random M, subject lengths to 64 rows, kernel structures, storages, panel shapes), fused likelihoods (random
variable layouts, 2..16 classes, float32 / float64 storage), observation heads (both backward kernels), batch
normalisation, GP prediction and the validation bound."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,cases", [("kl_stress.py", 16), ("loglik_stress.py", 16), ("theta_stress.py", 16),
                                        ("aux_stress.py", 8)])
def test_sweep(tool, cases):
    env = dict(os.environ, SEED="101", CASES=str(cases))
    for k in ("HLVAE_PANEL_RP", "HLVAE_PANEL_WAVES", "HLVAE_B200_LIB"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "tools", tool)], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0 and "stress: OK" in r.stdout, tail
