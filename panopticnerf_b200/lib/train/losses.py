"""Loss side of the training path (the reference's NetworkWrapper computes these terms with torch
ops on the rendered maps): photometric, depth, 2D pseudo-label cross-entropy on the rendered semantics and the
cross-entropy of the fixed (bounding-primitive) semantics - values and map gradients from ONE kernel (`pnr_losses`),
exposed as an autograd node so that `loss.backward()` continues into `raw2outputs_autograd` (pnr_composite_backward)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from ... import _capi

_MAPS = ("rgb_map", "rgb_map0", "depth_map", "semantic_map", "fixed_semantic_map")


def _run(maps: Dict[str, Optional[torch.Tensor]], rgb_gt, depth_gt, label, label_weight, w, sem_is_prob: bool, eps: float):
    ref = next(t for t in maps.values() if t is not None)
    dev, R = ref.device, ref.shape[0]
    keep = {k: (None if t is None else t.detach().to(torch.float32).contiguous()) for k, t in maps.items()}
    tg = {"rgb_gt": rgb_gt, "depth_gt": depth_gt, "label_weight": label_weight}
    keep.update({k: (None if t is None else t.detach().to(dev, torch.float32).contiguous()) for k, t in tg.items()})
    lab = None if label is None else label.detach().to(dev, torch.int32).contiguous()
    Cn = keep["semantic_map"].shape[1] if keep["semantic_map"] is not None else (
        keep["fixed_semantic_map"].shape[1] if keep["fixed_semantic_map"] is not None else 0)
    n_depth = int((keep["depth_gt"] > 0).sum()) if keep["depth_gt"] is not None else 0
    n_sem = int(((lab >= 0) & (lab < Cn)).sum()) if lab is not None else 0
    n_rgb = 3 * R
    a = _capi.PnrLossArgs()
    a.R, a.C, a.sem_is_prob = R, Cn, int(bool(sem_is_prob))
    for k in _MAPS + ("rgb_gt", "depth_gt", "label_weight"):
        setattr(a, k, _capi.ptr(keep[k], torch.float32, k) if keep[k] is not None else None)
    a.label = _capi.ptr(lab) if lab is not None else None
    a.w_rgb, a.w_depth, a.w_sem, a.w_fix = [float(x) for x in w]
    a.inv_n_rgb, a.inv_n_depth, a.inv_n_sem = 1.0 / n_rgb, 1.0 / max(n_depth, 1), 1.0 / max(n_sem, 1)
    a.eps = float(eps)
    per_ray = torch.empty(R, 4, dtype=torch.float32, device=dev)
    grads = {k: (torch.empty_like(keep[k]) if keep[k] is not None else None) for k in _MAPS}
    a.per_ray = _capi.ptr(per_ray)
    for k in _MAPS:
        setattr(a, "d_" + k, _capi.ptr(grads[k]) if grads[k] is not None else None)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().pnr_losses(C.byref(a), _capi.stream_ptr()), "pnr_losses")
    sums = per_ray.sum(0)
    terms = torch.stack([sums[0] * a.inv_n_rgb, sums[1] * a.inv_n_depth, sums[2] * a.inv_n_sem, sums[3] * a.inv_n_sem])
    total = (terms * torch.tensor([float(x) for x in w], device=dev)).sum()
    return total, terms, grads


class PanopticLoss(torch.autograd.Function):
    """total, terms = PanopticLoss.apply(rgb_map, rgb_map0, depth_map, semantic_map, fixed_semantic_map, rgb_gt,
    depth_gt, label, label_weight, (w_rgb, w_depth, w_sem, w_fix), sem_is_prob, eps).  Maps may be None; `terms`
    (the four means, unweighted) is not differentiable."""

    @staticmethod
    def forward(ctx, rgb_map, rgb_map0, depth_map, semantic_map, fixed_semantic_map, rgb_gt, depth_gt, label,
                label_weight, weights, sem_is_prob, eps):
        maps = dict(zip(_MAPS, (rgb_map, rgb_map0, depth_map, semantic_map, fixed_semantic_map)))
        total, terms, grads = _run(maps, rgb_gt, depth_gt, label, label_weight, weights, sem_is_prob, eps)
        ctx.grads = [grads[k] for k in _MAPS]
        ctx.mark_non_differentiable(terms)
        return total, terms

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        return tuple(None if g is None else g * g_total for g in ctx.grads) + (None,) * 7


def panoptic_losses(out: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor],
                    weights: Tuple[float, float, float, float] = (1.0, 0.1, 1.0, 1.0), sem_is_prob: bool = False,
                    eps: float = 1e-8, out_coarse: Optional[Dict[str, torch.Tensor]] = None):
    """Convenience wrapper over a Renderer result: batch keys rgb (gt) [R,3], depth (gt, <= 0 = invalid) [R],
    pseudo_label [R] int (-1 = ignore), pseudo_weight [R] (optional).  Returns (total, {'rgb','depth','sem','fix'})."""
    g = out.get
    total, terms = PanopticLoss.apply(g("rgb_map"), None if out_coarse is None else out_coarse.get("rgb_map"),
                                      g("depth_map") if "depth" in batch else None,
                                      g("semantic_map") if "pseudo_label" in batch else None,
                                      g("fixed_semantic_map") if "pseudo_label" in batch else None,
                                      batch.get("rgb"), batch.get("depth"), batch.get("pseudo_label"),
                                      batch.get("pseudo_weight"), tuple(weights), sem_is_prob, eps)
    return total, dict(zip(("rgb", "depth", "sem", "fix"), terms.unbind(0)))
