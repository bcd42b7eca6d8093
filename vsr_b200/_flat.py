"""The per-step check that a net's parameters still are views of its flat bucket (`net.flat`).

Every fused step asks twice (the step and FlatAdam).  Walking the module tree costs 100-200 us for the nets here - host time
the device idles through when the caller reads the loss back every step - so the walk is repeated only after SOME module
registered a parameter (`module.weight = nn.Parameter(...)`, `load_state_dict(assign=True)`: a global registration hook bumps
an epoch); otherwise the Parameter objects are the ones the last walk saw and only their storage pointers are compared
(~15 us for 84 parameters).  A new `net.flat` (every `_flatten()`) drops the cache by identity."""
import torch.nn as nn

_PARAM_EPOCH = [0]


def _on_parameter_registration(module, name, param):
    _PARAM_EPOCH[0] += 1
    return None


nn.modules.module.register_module_parameter_registration_hook(_on_parameter_registration)


def is_flat(net):
    flat = net.flat
    c = net.__dict__.get("_flat_cache")
    if c is not None and c[0] == _PARAM_EPOCH[0] and c[1] is flat:
        for p, e in zip(c[2], c[3]):
            if p.data_ptr() != e:
                return False
        return True
    base, es = flat.data_ptr(), flat.element_size()
    params, ptrs = [], []
    for p, ref in zip(net.parameters(), net._plan.params.values()):
        if p.data_ptr() != base + ref.offset * es or p.device != flat.device:
            return False
        params.append(p)
        ptrs.append(base + ref.offset * es)
    net.__dict__["_flat_cache"] = (_PARAM_EPOCH[0], flat, params, ptrs)
    return True
