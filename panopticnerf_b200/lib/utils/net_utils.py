"""Checkpoint reader for reference-trained weights (reference: lib/utils/net_utils.py
`load_network` / `load_model`, not in the mount - layout recalled from the zju3dv training template):

    <model_dir>/<epoch>.pth or latest.pth  ->  {'net': state_dict, 'optim': ..., 'scheduler': ..., 'recorder': ..., 'epoch': int}

where the `net` keys may carry a wrapper prefix (`net.`, `module.`, `network.`) because the trainer saves
the NetworkWrapper / DDP module.  Only the `net` entry is read; parameter names inside follow nerf-pytorch
(`pts_linears.0.weight` ...), which is also what panopticnerf_b200's Network uses, so loading is a key-prefix
strip plus `load_state_dict`.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional

import torch

_PREFIXES = ("module.", "net.", "network.", "model.")


def _strip(key: str) -> str:
    changed = True
    while changed:
        changed = False
        for p in _PREFIXES:
            if key.startswith(p):
                key, changed = key[len(p):], True
    return key


def find_checkpoint(model_dir: str, epoch: int = -1) -> Optional[str]:
    """`epoch` -1 -> latest.pth if present, else the highest numbered <epoch>.pth (the template's rule)."""
    if os.path.isfile(model_dir):
        return model_dir
    if not os.path.isdir(model_dir):
        return None
    if epoch >= 0:
        p = os.path.join(model_dir, f"{epoch}.pth")
        return p if os.path.exists(p) else None
    if os.path.exists(os.path.join(model_dir, "latest.pth")):
        return os.path.join(model_dir, "latest.pth")
    nums = [int(m.group(1)) for f in os.listdir(model_dir) if (m := re.fullmatch(r"(\d+)\.pth", f))]
    return os.path.join(model_dir, f"{max(nums)}.pth") if nums else None


def extract_state_dict(ckpt) -> Dict[str, torch.Tensor]:
    sd = ckpt["net"] if isinstance(ckpt, dict) and "net" in ckpt else ckpt
    if not isinstance(sd, dict):
        raise ValueError("checkpoint has no 'net' state_dict")
    return {_strip(k): v for k, v in sd.items()}


def load_network(net: torch.nn.Module, model_dir: str, epoch: int = -1, strict: bool = True,
                 only: Optional[str] = None) -> int:
    """Load reference weights into `net`.  `only` selects a sub-network by prefix when one checkpoint holds
    several (e.g. 'fine_net.').  Returns the stored epoch (0 if absent).  Raises if nothing is found."""
    path = find_checkpoint(model_dir, epoch)
    if path is None:
        raise FileNotFoundError(f"no checkpoint under {model_dir!r} (epoch={epoch})")
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = extract_state_dict(ckpt)
    if only is not None:
        sd = {_strip(k[len(only):]): v for k, v in sd.items() if k.startswith(only)}
    own = net.state_dict()
    if not strict:
        sd = {k: v for k, v in sd.items() if k in own and tuple(v.shape) == tuple(own[k].shape)}
    missing = [k for k in own if k not in sd]
    if strict and missing:
        raise KeyError(f"{path}: missing keys {missing[:4]}{'...' if len(missing) > 4 else ''}")
    net.load_state_dict({k: sd[k] for k in own if k in sd}, strict=strict)
    return int(ckpt.get("epoch", 0)) if isinstance(ckpt, dict) else 0
