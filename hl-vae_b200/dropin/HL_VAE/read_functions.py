"""Drop-in for HL_VAE/read_functions.py: the reference module (data reading, error metrics, ...)
with `statistics` (:268-339) and `discrete_variables_transformation` (:221-235) replaced by the
CUDA kernels for the five hot-path variable types.  Types outside that set (beta) fall through to
the reference code."""
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


def _supported(types_info):
    return all(t['type'] in ('real', 'pos', 'count', 'cat', 'ordinal') for t in types_info['types_dict'])


def statistics(loglik_params, types_info, device, conv=False, log_vy=None):
    if not _supported(types_info) or not loglik_params.is_cuda:
        if _ref is None:
            raise RuntimeError("hlvae_b200: statistics needs CUDA tensors (reference module not found for other cases)")
        return _ref.statistics(loglik_params, types_info, device, conv, log_vy)
    lay = _layout(types_info, loglik_params.device)
    lv_pos = None
    if log_vy is not None and lay.idx["pos"].numel():
        lv_pos = log_vy[1]
    vparam = torch.zeros(4, lay.D, dtype=torch.float64, device=loglik_params.device)
    if lv_pos is not None:
        vparam[2, lay.idx["pos"]] = lv_pos.detach().to(torch.float64)[lay.gpos["pos"]]
    return _ll.statistics(lay, loglik_params, vparam)


def discrete_variables_transformation(data, types_info):
    if not _supported(types_info) or not data.is_cuda:
        if _ref is None:
            raise RuntimeError("hlvae_b200: discrete_variables_transformation needs CUDA tensors")
        return _ref.discrete_variables_transformation(data, types_info)
    return _ll.discrete_variables_transformation(_layout(types_info, data.device), data)
