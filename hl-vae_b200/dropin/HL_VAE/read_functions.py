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

_layouts = {}


def _layout(types_info, device):
    key = (id(types_info), str(device))
    if key not in _layouts:
        _layouts[key] = _ll.VarLayout.from_types_info(types_info, device)
    return _layouts[key]


def _require(types_info, tensor, what):
    bad = sorted({t['type'] for t in types_info['types_dict']} - set(_ll.SUPPORTED_TYPES))
    if bad:
        raise NotImplementedError(f"hlvae_b200: {what}: variable types {bad} are not supported by the CUDA kernels")
    if not tensor.is_cuda:
        raise RuntimeError(f"hlvae_b200: {what} runs on CUDA tensors only (no CPU fallback)")


def statistics(loglik_params, types_info, device, conv=False, log_vy=None):
    _require(types_info, loglik_params, "statistics")
    lay = _layout(types_info, loglik_params.device)
    lv_pos = None
    if log_vy is not None and lay.idx["pos"].numel():
        lv_pos = log_vy[1]
    vparam = torch.zeros(4, lay.D, dtype=torch.float64, device=loglik_params.device)
    if lv_pos is not None:
        vparam[2, lay.idx["pos"]] = lv_pos.detach().to(torch.float64)[lay.gpos["pos"]]
    return _ll.statistics(lay, loglik_params, vparam)


def discrete_variables_transformation(data, types_info):
    _require(types_info, data, "discrete_variables_transformation")
    return _ll.discrete_variables_transformation(_layout(types_info, data.device), data)
