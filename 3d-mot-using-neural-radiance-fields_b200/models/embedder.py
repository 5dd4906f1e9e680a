"""Positional encoding front-end (models/embedder.py:26-112 of the reference).

`Embedder.forward` runs the stand-alone CUDA kernel (star_embed); inside NeRF.forward the encoding
is computed in registers by the fused MLP kernel and only the BARF mask vector comes from here."""
import math

import torch
from torch import nn

from .. import functional as F_


def barf_weights(step, end_barf, L, start_barf=0):
    """w_k = (1 - cos(pi * clamp(alpha - k, 0, 1))) / 2,  alpha = (step-start)/(end-start) * L  (:26-30)."""
    alpha = (step - start_barf) / (end_barf - start_barf) * L
    k = torch.arange(L, dtype=torch.float32)
    return (1 - (alpha - k).clamp(min=0, max=1).mul(math.pi).cos()) / 2


def barf_scale_vector(step, end_barf, L, d=3, pad_to=None):
    """Per-element mask of the [x, sin, cos, ...] encoding.  Reference quirk (:32,109): the mask is
    applied as enc[:, d:].view(-1, L) * w, so encoded element j (0-based after the raw input) is
    scaled by w[j mod L], not by the weight of its own frequency."""
    n = d + 2 * d * L
    s = torch.ones(pad_to or n, dtype=torch.float32)
    w = barf_weights(step, end_barf, L)
    j = torch.arange(2 * d * L)
    s[d:n] = w[j % L]
    return s


class Embedder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        if kwargs.get("input_dims", 3) != 3 or not kwargs.get("include_input", True) or \
                not kwargs.get("log_sampling", True):
            raise NotImplementedError("B200 path encodes 3-D inputs with include_input and log sampling only")
        self.L = int(kwargs["num_freqs"])
        self.out_dim = 3 + 6 * self.L
        self._scale_cache = {}

    def scale(self, step, device, pad_to=None):
        """BARF mask on `device`, or None when BARF is off (step None or end_barf == -1, :99)."""
        end_barf = self.kwargs["end_barf"]
        if step is None or end_barf == -1:
            return None
        key = (float(step), str(device), pad_to)
        if key not in self._scale_cache:
            if len(self._scale_cache) > 64:
                self._scale_cache.clear()
            self._scale_cache[key] = barf_scale_vector(step, end_barf, self.L, 3, pad_to).to(device)
        return self._scale_cache[key]

    def forward(self, inputs, step=None):
        return F_.embed(inputs, self.L, self.scale(step, inputs.device))


def get_embedder(multires, end_barf, i=0, input_dims=3):
    if i == -1:
        return nn.Identity(), 3
    emb = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                   log_sampling=True, periodic_fns=[torch.sin, torch.cos], end_barf=end_barf)
    return emb, emb.out_dim
