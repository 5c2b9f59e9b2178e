"""CPU oracle for the steps after / beside the render path.  TEST INFRASTRUCTURE ONLY (same rule as
reference_renderer.py: only tests/, smoke() and bench.py's CPU legs may import it).

PARITY UNPINNED: the reference's panoptic fusion and its 360 branch's hash-grid encoder are not in the mount
(/root/reference/README.md:7,13 point at the code branches).  `panoptic_fuse` states the fusion rule chosen in this
repo; `hashgrid_encode` restates the published multi-resolution hash encoding (Mueller et al. 2022, section 3:
N_l = floor(N_min b^l), dense indexing while (N_l+1)^3 <= T, else the spatial hash with primes 1, 2654435761,
805459861, trilinear interpolation) in fp32 with a fixed operation order."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch


def _argmax_first(v: torch.Tensor, keep: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Row-wise argmax with NaN = -inf and ties -> lowest index; -1 where no channel is kept."""
    x = torch.where(torch.isnan(v), torch.full_like(v, -math.inf), v)
    if keep is not None:
        x = torch.where(keep, x, torch.full_like(x, -math.inf))
        none = ~keep.any(dim=1)
    else:
        none = torch.zeros(v.shape[0], dtype=torch.bool)
    best = x.max(dim=1, keepdim=True).values
    cand = (x == best) if keep is None else ((x == best) & keep)
    idx = torch.where(cand, torch.arange(v.shape[1])[None].expand_as(cand), torch.full_like(cand, v.shape[1], dtype=torch.long)).min(dim=1).values
    return torch.where(none, torch.full_like(idx, -1), idx)


def panoptic_fuse(sem: torch.Tensor, inst: Optional[torch.Tensor], is_thing, inst_class, inst_id=None, class_id=None,
                  palette=None):
    """Returns (panoptic int32 [R], sem_label int16, inst_slot int16, color uint8 [R,3]).  Rule: include/pnr.h."""
    R, C = sem.shape
    s = _argmax_first(sem)
    k = torch.full((R,), -1, dtype=torch.long)
    if inst is not None and inst.shape[1] > 0:
        ic = torch.as_tensor(inst_class, dtype=torch.long)
        thing = torch.as_tensor(is_thing, dtype=torch.bool)[s.clamp_min(0)] & (s >= 0)
        keep = (ic[None, :] == s[:, None]) & thing[:, None]
        k = _argmax_first(inst, keep)
    cid = s if class_id is None else torch.as_tensor(class_id, dtype=torch.long)[s.clamp_min(0)]
    stuff = cid * 1000
    if inst_id is None:
        thing_id = cid * 1000 + k + 1
    else:
        thing_id = torch.as_tensor(inst_id, dtype=torch.long)[k.clamp_min(0)]
    pan = torch.where(k >= 0, thing_id, stuff)
    pan = torch.where(s >= 0, pan, torch.full_like(pan, -1))
    col = torch.zeros(R, 3, dtype=torch.long)
    if palette is not None:
        col = torch.as_tensor(palette, dtype=torch.long).reshape(-1, 3)[s.clamp_min(0)] * (s >= 0)[:, None]
    h = (pan.to(torch.int64) & 0xFFFFFFFF) * 2654435761 & 0xFFFFFFFF
    hc = torch.stack([(h >> 8) & 0xFF, (h >> 16) & 0xFF, (h >> 24) & 0xFF], -1)
    col = torch.where((k >= 0)[:, None], (col + hc + 1) >> 1, col)
    return pan.to(torch.int32), s.to(torch.int16), k.to(torch.int16), col.to(torch.uint8)


def hashgrid_resolutions(L: int, base: float, scale: float):
    return [int(math.floor(float(np.float32(base)) * float(np.float32(scale)) ** l)) for l in range(L)]


def hashgrid_encode(x: torch.Tensor, aabb: Optional[torch.Tensor], table: torch.Tensor, base: float, scale: float) -> torch.Tensor:
    """x [n,3] fp32, aabb [2,3] or None, table [L,T,F] fp32 (T a power of two) -> [n, L*F] fp32."""
    L, T, F = table.shape
    v = x.to(torch.float32)
    if aabb is not None:
        a = aabb.to(torch.float32).reshape(2, 3)
        v = (v - a[0]) / (a[1] - a[0])
    v = torch.clamp(v, 0.0, 1.0)
    outs = []
    for l, res in enumerate(hashgrid_resolutions(L, base, scale)):
        res_f = torch.tensor(float(res), dtype=torch.float32)
        p = v * res_f
        fl = torch.floor(p)
        fl = torch.where(fl >= res_f, res_f - 1.0, fl)
        w = p - fl
        c = fl.to(torch.int64)
        res1 = res + 1
        dense = res1 ** 3 <= T
        acc = torch.zeros(v.shape[0], F, dtype=torch.float32)
        for k in range(8):
            dx, dy, dz = k & 1, (k >> 1) & 1, k >> 2
            wx = w[:, 0] if dx else 1.0 - w[:, 0]
            wy = w[:, 1] if dy else 1.0 - w[:, 1]
            wz = w[:, 2] if dz else 1.0 - w[:, 2]
            wk = (wx * wy) * wz
            cx, cy, cz = c[:, 0] + dx, c[:, 1] + dy, c[:, 2] + dz
            if dense:
                idx = cx + cy * res1 + cz * res1 * res1
            else:
                idx = ((cx & 0xFFFFFFFF) ^ ((cy * 2654435761) & 0xFFFFFFFF) ^ ((cz * 805459861) & 0xFFFFFFFF)) & (T - 1)
            acc = acc + wk[:, None] * table[l][idx]
        outs.append(acc)
    return torch.cat(outs, -1)
