"""Deterministic synthetic weights (there is no checkpoint offline).

``synthetic_state_dict(cfg, seed)`` fills every entry of the model's ``state_dict`` from a seeded CPU generator in
key order, so that the product, the oracle and the reference model (in ``oracle/make_golden.py``) all run with the
same numbers on any machine of this image.  Scales follow the PyTorch defaults (U(-1/sqrt(fan_in), 1/sqrt(fan_in)));
norm layers get a non-trivial affine so that it is exercised.
"""
import math

import torch

from .modules.kpconv import default_kernel_points


def synthetic_state_dict(model, seed=7351):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    kp_seed = seed
    radii = {}
    for name, mod in model.named_modules():
        if mod.__class__.__name__ == 'KPConv':
            radii[name + '.kernel_points'] = (mod.kernel_size, mod.radius)
    for name, t in model.state_dict().items():
        shape = tuple(t.shape)
        if name.endswith('kernel_points'):
            kp_seed += 1
            k, r = radii[name]
            v = default_kernel_points(k, r, seed=kp_seed)
        elif name.endswith('div_term'):
            v = t.clone()
        elif name.endswith('alpha'):
            v = torch.tensor(1.0)
        elif '.norm.' in name or name.endswith('norm.weight') or name.endswith('norm.bias'):
            if name.endswith('weight'):
                v = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                v = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith('KPConv.weights'):
            bound = 1.0 / math.sqrt(shape[1] * shape[2]) * 4.0
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith('weight') and len(shape) == 2:
            bound = 1.0 / math.sqrt(shape[1])
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith('bias'):
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        else:
            raise KeyError(f'no init rule for {name} {shape}')
        sd[name] = v.float().contiguous()
    return sd
