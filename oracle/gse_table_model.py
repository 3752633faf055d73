"""TEST INFRASTRUCTURE (never imported by the product): numpy restatement of the arithmetic of csrc/gse_table.cu -- the tabulated
form of GeometricStructureEmbedding.forward (/root/reference/geotransformer/modules/geotransformer/geotransformer.py:57-72) --
so that the accuracy of the tabulation itself (grid step, fp16 forward differences, one power-of-two scale) can be checked
against the oracle on the CPU.  Table: node i holds fp32(g(i h)) and fp16((g((i + 1) h) - g(i h)) / scale), built in fp64;
lookup: t = x / h, i = floor(t), value = fma((t - i) * scale, diff_i, value_i) in fp32."""
import math

import numpy as np
import torch

from . import geo_oracle as G


def _build(weight, bias, div, x_max, inv_step):
    n = int(math.ceil(x_max * inv_step)) + 1
    x0 = np.arange(n, dtype=np.float64) / inv_step
    w = weight.double().numpy()

    def g(x):
        om = x[:, None] * div[None, :]
        s = np.stack([np.sin(om), np.cos(om)], axis=2).reshape(len(x), 2 * len(div))
        return s @ w.T

    a0, a1 = g(x0), g(x0 + 1.0 / inv_step)
    return (a0 + bias.double().numpy()).astype(np.float32), a1 - a0


def structure_embedding_tabulated(sd, pre, points, sigma_d, sigma_a, angle_k, inv_step=256, d_max=96.0):
    div = sd[pre + 'embedding.div_term'].double().numpy()
    d_idx, a_idx = G.embedding_indices(points, sigma_d, sigma_a, angle_k)
    wd, wa = sd[pre + 'proj_d.weight'], sd[pre + 'proj_a.weight']
    vd, dd = _build(wd, sd[pre + 'proj_d.bias'], div, d_max, inv_step)
    va, da = _build(wa, sd[pre + 'proj_a.bias'], div, 180.0 / sigma_a + 0.25, inv_step)
    freq = np.repeat(div, 2)
    bound = max((np.abs(wd.double().numpy()) * freq).sum(1).max(), (np.abs(wa.double().numpy()) * freq).sum(1).max()) / inv_step
    scale = np.float32(2.0 ** (math.frexp(float(np.float32(bound)))[1] - 14))
    hd = (dd / scale).astype(np.float32).astype(np.float16).astype(np.float32)
    ha = (da / scale).astype(np.float32).astype(np.float16).astype(np.float32)

    def look(values, diffs, x):
        t = x.astype(np.float32) * np.float32(inv_step)
        i = t.astype(np.int64)
        assert int(i.max()) < len(values), 'argument beyond the table (the kernel evaluates those directly)'
        fr = (t - i.astype(np.float32)) * scale
        return (fr[..., None] * diffs[i] + values[i]).astype(np.float32)

    e = look(vd, hd, d_idx.numpy()) + look(va, ha, a_idx.numpy()).max(axis=2)
    return torch.from_numpy(e)
