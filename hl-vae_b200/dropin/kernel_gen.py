"""Drop-in for the reference's kernel_gen module (HLVAE_main.py:15 imports generate_kernel_batched).
The unbatched legacy builders generate_kernel / generate_kernel_approx (kernel_gen.py:9-197, unused by
HLVAE_main.py) are passed through from the reference module when it is on sys.path."""
import importlib.util
import os
import sys

from hlvae_b200.kernels import generate_kernel_batched  # noqa: F401

_here = os.path.dirname(os.path.abspath(__file__))
for _p in sys.path:
    _f = os.path.join(_p or ".", "kernel_gen.py")
    if os.path.exists(_f) and os.path.dirname(os.path.abspath(_f)) != _here:
        try:
            _spec = importlib.util.spec_from_file_location("_hlvae_reference_kernel_gen", _f)
            _ref = importlib.util.module_from_spec(_spec)
            _spec.loader.exec_module(_ref)
            for _k in ("generate_kernel", "generate_kernel_approx"):
                if hasattr(_ref, _k):
                    globals().setdefault(_k, getattr(_ref, _k))
        except Exception:      # the legacy builders need gpytorch; the batched one served here does not
            pass
        break
