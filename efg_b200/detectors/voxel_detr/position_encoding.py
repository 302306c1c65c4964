"""Sine position embedding over the BEV map (VD/modules/position_encoding.py:7-62), normalised,
temperature 10000, scale 2*pi.  Depends only on the map size, so it is cached per (H, W, device)."""
import math

import torch
from torch import nn


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = {}

    def _build(self, h, w, dtype, device):
        y = torch.arange(1, h + 1, dtype=dtype, device=device)[:, None].expand(h, w)
        x = torch.arange(1, w + 1, dtype=dtype, device=device)[None, :].expand(h, w)
        if self.normalize:
            eps = 1e-6
            y = (y - 0.5) / (y[-1:, :] + eps) * self.scale
            x = (x - 0.5) / (x[:, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * dim_t.div(2, rounding_mode="floor") / self.num_pos_feats)
        px = x[:, :, None] / dim_t
        py = y[:, :, None] / dim_t
        px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
        py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
        return torch.cat((py, px), dim=2).permute(2, 0, 1).contiguous()  # [2*F, H, W]

    def forward(self, x, mask=None):
        assert mask is None, "masked BEV maps are not produced on this path"
        h, w = x.shape[-2:]
        key = (h, w, x.dtype, x.device)
        if key not in self._cache:
            self._cache[key] = self._build(h, w, x.dtype, x.device)
        return self._cache[key][None].expand(x.shape[0], -1, -1, -1)


def build_position_encoding(kind, hidden_dim):
    if kind in ("v2", "sine"):
        return PositionEmbeddingSine(hidden_dim // 2, normalize=True)
    raise ValueError("not supported {}".format(kind))
