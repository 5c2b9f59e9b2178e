"""Multi-resolution hash-grid feature encoder (the 360 model's extra encoder), forward only:
`pnr_hashgrid_encode`.  The table is an fp32 parameter [L, 2^T_log2, F]; gradients w.r.t. it are not built yet."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .... import _capi


class HashGrid(nn.Module):
    def __init__(self, n_levels: int = 16, n_features: int = 2, log2_hashmap_size: int = 19, base_resolution: float = 16.0,
                 per_level_scale: float = 1.3819, aabb: Optional[torch.Tensor] = None, seed: int = 0):
        super().__init__()
        self.L, self.F, self.T_log2 = int(n_levels), int(n_features), int(log2_hashmap_size)
        self.base, self.scale = float(base_resolution), float(per_level_scale)
        g = torch.Generator().manual_seed(seed)
        self.table = nn.Parameter((torch.rand(self.L, 1 << self.T_log2, self.F, generator=g) * 2 - 1) * 1e-4, requires_grad=False)
        self.register_buffer("aabb", None if aabb is None else torch.as_tensor(aabb, dtype=torch.float32).reshape(6).clone())

    @property
    def out_dim(self) -> int:
        return self.L * self.F

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        p = _capi.ptr
        xc = x.reshape(-1, 3).to(torch.float32).contiguous()
        xp, tp = p(xc, torch.float32, "x"), p(self.table, torch.float32, "table")     # CPU tensors raise here
        out = torch.empty(xc.shape[0], self.out_dim, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _capi.check(_capi.lib().pnr_hashgrid_encode(xp, xc.shape[0], p(self.aabb), tp, self.L, self.F, self.T_log2,
                                                        self.base, self.scale, p(out), _capi.stream_ptr()), "pnr_hashgrid_encode")
        return out.reshape(*x.shape[:-1], self.out_dim)
