"""Panoptic label fusion + colour mapping of a render's maps (the reference's evaluator /
visualiser does this on the CPU after the render).  One kernel, one warp per ray: `pnr_panoptic_fuse`."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from ... import _capi


def fuse_panoptic(out: Dict[str, torch.Tensor], is_thing: Sequence[int], inst_class: Optional[Sequence[int]] = None,
                  inst_id: Optional[Sequence[int]] = None, class_id: Optional[Sequence[int]] = None,
                  palette: Optional[torch.Tensor] = None, sem_key: str = "semantic_map",
                  inst_key: str = "instance_map") -> Dict[str, torch.Tensor]:
    """out: a Renderer.render result (CUDA).  Returns {'panoptic' i32 [R], 'semantic' i16, 'instance_slot' i16,
    'color' u8 [R,3] (when a palette is given)}.  Fusion rule: include/pnr.h (pnr_panoptic_fuse)."""
    sem = out[sem_key]
    inst = out.get(inst_key)
    dev = sem.device
    R, C = sem.shape
    K = inst.shape[1] if inst is not None else 0
    i32 = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.int32).to(dev).contiguous()
    thing = torch.as_tensor(is_thing, dtype=torch.uint8).to(dev).contiguous()
    if thing.numel() != C:
        raise ValueError(f"fuse_panoptic: is_thing has {thing.numel()} entries for {C} classes")
    ic, ii, ci = i32(inst_class), i32(inst_id), i32(class_id)
    if K > 0 and (ic is None or ic.numel() != K):
        raise ValueError(f"fuse_panoptic: inst_class must list the class of each of the {K} instance slots")
    pal = None if palette is None else torch.as_tensor(palette, dtype=torch.uint8).to(dev).reshape(C, 3).contiguous()
    res = {"panoptic": torch.empty(R, dtype=torch.int32, device=dev), "semantic": torch.empty(R, dtype=torch.int16, device=dev),
           "instance_slot": torch.empty(R, dtype=torch.int16, device=dev)}
    if pal is not None:
        res["color"] = torch.empty(R, 3, dtype=torch.uint8, device=dev)
    p = _capi.ptr
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().pnr_panoptic_fuse(p(sem, torch.float32, sem_key), p(inst, torch.float32, inst_key) if K else None,
                                                  R, C, K, p(thing), p(ic), p(ii), p(ci), p(pal), p(res["panoptic"]),
                                                  p(res["semantic"]), p(res["instance_slot"]), p(res.get("color")),
                                                  _capi.stream_ptr()), "pnr_panoptic_fuse")
    return res
