"""Drop-in for HL_VAE/read_functions.py: the reference module (data reading, error metrics, ...)
with `statistics` (:268-339) and `discrete_variables_transformation` (:221-235) replaced by the
CUDA kernels.  There is no CPU path: CPU tensors and variable types the kernels do not know raise."""
import importlib.util
import os

import torch

from hlvae_b200 import loglik as _ll

_ref = None
for _p in __import__("HL_VAE").__path__[1:]:
    _f = os.path.join(_p, "read_functions.py")
    if os.path.exists(_f):
        _spec = importlib.util.spec_from_file_location("_hlvae_reference_read_functions", _f)
        _ref = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_ref)
        globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
        break

_layouts = {}          # content-keyed (types, device), see hlvae_b200.normalize._layout_for


def _require_cuda(tensor, what):
    if not tensor.is_cuda:
        raise RuntimeError(f"hlvae_b200: {what} runs on CUDA tensors only (no CPU fallback)")


def statistics(loglik_params, types_info, device, conv=False, log_vy=None):
    _require_cuda(loglik_params, "statistics")
    return _ll.statistics_general(loglik_params, types_info, conv, log_vy)


def discrete_variables_transformation(data, types_info):
    _require_cuda(data, "discrete_variables_transformation")
    types = [("real" if t['type'] == "beta" else t['type'], int(t['nclass'])) for t in types_info['types_dict']]
    key = (tuple(types), str(data.device))
    if key not in _layouts:
        _layouts[key] = _ll.VarLayout(types, data.device)
    return _ll.discrete_variables_transformation(_layouts[key], data)
