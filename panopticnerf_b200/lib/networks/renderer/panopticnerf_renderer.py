"""Renderer / raw2outputs / sample_pdf — the volume-render loop of PanopticNeRF (SURVEY.md 8(a)
a3, a4, a5, a6, a9, a10; reference: the Renderer module under lib/networks/renderer/ and its
raw2outputs / sample_pdf helpers, not in the mount).

Same names, argument meaning and returned keys as the oracle restatement (oracle/reference_renderer.py),
but every stage is a libpnr CUDA kernel called through the C ABI; torch only owns the device buffers.
CPU tensors raise: there is no CPU path in the product.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Dict, Optional

import torch

from panopticnerf_b200 import _capi

_F32, _I32 = torch.float32, torch.int32
MAX_SAMPLES_PER_RAY = 256      # pnr_composite: 32 lanes x 8 samples


def _f(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _capi.PnrError(f"{name}: expected a CUDA tensor, got {t.device} (panopticnerf_b200 is GPU-only)")
    return t.to(_F32).contiguous()


def _on_tensor_device(fn):
    """Run a stage wrapper with the device of its first tensor argument current: the stream handed to libpnr
    (torch's current stream) and the kernels it launches then belong to the device the buffers live on, whatever
    the caller's current device is (several GPUs driven from one process)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        t = next((a for a in args if torch.is_tensor(a)), None)
        if t is None or not t.is_cuda:
            return fn(*args, **kw)          # the wrapper's own checks raise the "GPU-only" error
        with torch.cuda.device(t.device):
            return fn(*args, **kw)
    return wrapper


def _same_device(ref: torch.Tensor, **tensors) -> None:
    for name, t in tensors.items():
        if t is not None and t.device != ref.device:
            raise _capi.PnrError(f"{name} is on {t.device}, expected {ref.device}")


# ------------------------------------------------------------------------------------------------
# stage wrappers (one libpnr call each)
# ------------------------------------------------------------------------------------------------
@_on_tensor_device
def intersect(rays, box_center, box_half, box_rot, max_hits: int):
    """a5 -> hit_mask [R] bool, box_id [R,M] i32, t_in, t_out [R,M]."""
    rays = _f(rays, "rays")
    R, B, M = rays.shape[0], box_center.shape[0], int(max_hits)
    dev = rays.device
    hit = torch.empty(R, dtype=torch.uint8, device=dev)
    box_id = torch.empty(R, M, dtype=_I32, device=dev)
    t_in = torch.empty(R, M, dtype=_F32, device=dev)
    t_out = torch.empty(R, M, dtype=_F32, device=dev)
    bc, bh, br = _f(box_center, "box_center"), _f(box_half, "box_half"), _f(box_rot, "box_rot")
    _capi.check(_capi.lib().pnr_intersect(_capi.ptr(rays), R, _capi.ptr(bc), _capi.ptr(bh), _capi.ptr(br), B, M,
                                          _capi.ptr(hit), _capi.ptr(box_id), _capi.ptr(t_in), _capi.ptr(t_out),
                                          _capi.stream_ptr()), "pnr_intersect")
    return hit.bool(), box_id, t_in, t_out


@_on_tensor_device
def scene_near_far(rays, aabb, near_min: float, far_default: float):
    rays = _f(rays, "rays")
    R = rays.shape[0]
    near = torch.empty(R, dtype=_F32, device=rays.device)
    far = torch.empty(R, dtype=_F32, device=rays.device)
    a = (C.c_float * 6)(*[float(x) for x in aabb.detach().to("cpu", _F32).reshape(-1).tolist()])
    _capi.check(_capi.lib().pnr_scene_near_far(_capi.ptr(rays), R, a, float(near_min), float(far_default),
                                               _capi.ptr(near), _capi.ptr(far), _capi.stream_ptr()),
                "pnr_scene_near_far")
    return near, far


@_on_tensor_device
def bound_by_primitives(hit, box_id, t_in, t_out, near, far):
    hit8 = hit.to(torch.uint8).contiguous()
    near, far = near.clone(), far.clone()
    _capi.check(_capi.lib().pnr_bound_by_primitives(_capi.ptr(hit8), _capi.ptr(box_id, _I32), _capi.ptr(t_in),
                                                    _capi.ptr(t_out), near.shape[0], box_id.shape[1],
                                                    _capi.ptr(near), _capi.ptr(far), _capi.stream_ptr()),
                "pnr_bound_by_primitives")
    return near, far


@_on_tensor_device
def stratified_z(near, far, t_vals, perturb: float = 0.0, u: Optional[torch.Tensor] = None,
                 box_id=None, t_in=None, t_out=None, want_tags: bool = False):
    """a6 -> z [R,N] (and sample_box [R,N] i32 when want_tags)."""
    near, far, t_vals = _f(near, "near"), _f(far, "far"), _f(t_vals, "t_vals")
    R, N = near.shape[0], t_vals.shape[0]
    if perturb > 0.0 and u is None:
        u = torch.rand(R, N, device=near.device, dtype=_F32)
    u = _f(u, "u") if (u is not None and perturb > 0.0) else None
    z = torch.empty(R, N, dtype=_F32, device=near.device)
    sb = torch.empty(R, N, dtype=_I32, device=near.device) if want_tags else None
    M = box_id.shape[1] if box_id is not None else 0
    _capi.check(_capi.lib().pnr_sample_stratified(
        _capi.ptr(near), _capi.ptr(far), _capi.ptr(t_vals), _capi.ptr(u), R, N, float(perturb),
        _capi.ptr(box_id, _I32) if box_id is not None else None, _capi.ptr(t_in), _capi.ptr(t_out), M,
        _capi.ptr(z), _capi.ptr(sb), _capi.stream_ptr()), "pnr_sample_stratified")
    return (z, sb) if want_tags else z


@_on_tensor_device
def interval_z(near, far, t_vals, box_id, t_in, t_out, perturb: float = 0.0, u: Optional[torch.Tensor] = None):
    """a6 interval mode -> (z [R,N] ascending, sample_box [R,N] i32): the N samples sit inside the ray's hit
    intervals (n_m ~ N * len_m / sum len, remainder to the nearest; rays without a hit use the uniform rule)."""
    near, far, t_vals = _f(near, "near"), _f(far, "far"), _f(t_vals, "t_vals")
    R, N = near.shape[0], t_vals.shape[0]
    if perturb > 0.0 and u is None:
        u = torch.rand(R, N, device=near.device, dtype=_F32)
    u = _f(u, "u") if (u is not None and perturb > 0.0) else None
    z = torch.empty(R, N, dtype=_F32, device=near.device)
    sb = torch.empty(R, N, dtype=_I32, device=near.device)
    _capi.check(_capi.lib().pnr_sample_intervals(
        _capi.ptr(near), _capi.ptr(far), _capi.ptr(t_vals), _capi.ptr(u), R, N, float(perturb),
        _capi.ptr(box_id, _I32), _capi.ptr(t_in), _capi.ptr(t_out), box_id.shape[1], _capi.ptr(z), _capi.ptr(sb),
        _capi.stream_ptr()), "pnr_sample_intervals")
    return z, sb


@_on_tensor_device
def tag_samples(z, box_id, t_in, t_out):
    z = _f(z, "z")
    sb = torch.empty(z.shape, dtype=_I32, device=z.device)
    _capi.check(_capi.lib().pnr_tag_samples(_capi.ptr(z), z.shape[0], z.shape[1], _capi.ptr(box_id, _I32),
                                            _capi.ptr(t_in), _capi.ptr(t_out), box_id.shape[1], _capi.ptr(sb),
                                            _capi.stream_ptr()), "pnr_tag_samples")
    return sb


def generate_rays(H: int, W: int, intr, c2w: torch.Tensor, camera: str = "pinhole", row0: int = 0,
                  rows: Optional[int] = None, device="cuda") -> torch.Tensor:
    """Camera rays on the device: only 16 floats cross the host/device boundary."""
    rows = H - row0 if rows is None else rows
    rays = torch.empty(rows * W, 6, dtype=_F32, device=device)
    cam = {"pinhole": 0, "equirect": 1, "fisheye": 2}[camera]
    vals = [float(x) for x in intr]
    if len(vals) != (7 if cam == 2 else 4):
        raise ValueError(f"generate_rays: camera '{camera}' takes {7 if cam == 2 else 4} intrinsics, got {len(vals)}")
    k = (C.c_float * len(vals))(*vals)
    m = (C.c_float * 12)(*[float(x) for x in c2w.detach().to("cpu", _F32).reshape(-1)[:12].tolist()])
    with torch.cuda.device(rays.device):
        _capi.check(_capi.lib().pnr_generate_rays(int(H), int(W), int(row0), int(rows), cam,
                                                  k, m, _capi.ptr(rays), _capi.stream_ptr()), "pnr_generate_rays")
    return rays


@_on_tensor_device
def embed(x: torch.Tensor, L: int) -> torch.Tensor:
    """a7 standalone positional encoding (the Renderer uses the copy fused into the MLP kernel)."""
    xf = _f(x.reshape(-1, 3), "x")
    out = torch.empty(xf.shape[0], 3 + 6 * L, dtype=_F32, device=xf.device)
    _capi.check(_capi.lib().pnr_encode(_capi.ptr(xf), xf.shape[0], int(L), _capi.ptr(out), _capi.stream_ptr()),
                "pnr_encode")
    return out.reshape(*x.shape[:-1], 3 + 6 * L)


def _box_table_size(raw, sample_box, box_sem, box_inst) -> int:
    """One B bounds both id tables in the kernel (sample_box indexes them): they must have the same length and
    live on raw's device."""
    _same_device(raw, sample_box=sample_box, box_sem=box_sem, box_inst=box_inst)
    if sample_box is None:
        return 0
    sizes = {int(t.shape[0]) for t in (box_sem, box_inst) if t is not None}
    if len(sizes) > 1:
        raise ValueError(f"box_sem and box_inst must have one entry per primitive each, got lengths {sorted(sizes)}")
    return sizes.pop() if sizes else 0


@_on_tensor_device
def raw2outputs(raw, z_vals, rays_d, raw_noise_std: float = 0.0, white_bkgd: bool = False,
                num_classes: int = 0, num_instances: int = 0, sem_activation: str = "none",
                sample_box: Optional[torch.Tensor] = None, box_sem: Optional[torch.Tensor] = None,
                box_inst: Optional[torch.Tensor] = None, mask_outside: bool = False,
                noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """a9.  raw [R,N,4+C+K], z_vals [R,N], rays_d [R,3] (or rays [R,6])."""
    raw, z_vals = _f(raw, "raw"), _f(z_vals, "z_vals")
    R, N = z_vals.shape
    Cn, Kn = int(num_classes), int(num_instances)
    if raw.shape[-1] != 4 + Cn + Kn:
        raise ValueError(f"raw2outputs: raw has {raw.shape[-1]} channels, expected {4 + Cn + Kn}")
    if raw_noise_std > 0.0:
        nz = noise if noise is not None else torch.randn(R, N, device=raw.device)
        raw = raw.clone()
        raw[..., 3] += nz.to(raw.device, _F32) * raw_noise_std
    if rays_d.shape[-1] == 3:
        rays = torch.cat([torch.zeros_like(rays_d), rays_d], -1)
    else:
        rays = rays_d
    rays = _f(rays, "rays")
    dev = raw.device
    out = {"rgb_map": torch.empty(R, 3, dtype=_F32, device=dev), "depth_map": torch.empty(R, dtype=_F32, device=dev),
           "acc_map": torch.empty(R, dtype=_F32, device=dev), "disp_map": torch.empty(R, dtype=_F32, device=dev),
           "weights": torch.empty(R, N, dtype=_F32, device=dev)}
    if Cn > 0:
        out["semantic_map"] = torch.empty(R, Cn, dtype=_F32, device=dev)
    if Kn > 0:
        out["instance_map"] = torch.empty(R, Kn, dtype=_F32, device=dev)
    B = _box_table_size(raw, sample_box, box_sem, box_inst)
    if sample_box is not None and box_sem is not None and Cn > 0:
        out["fixed_semantic_map"] = torch.empty(R, Cn, dtype=_F32, device=dev)
    if sample_box is not None and box_inst is not None and Kn > 0:
        out["fixed_instance_map"] = torch.empty(R, Kn, dtype=_F32, device=dev)
    co = _capi.PnrCompositeOut(**{k: _capi.ptr(out[k]) if k in out else None
                                  for k, _ in _capi.PnrCompositeOut._fields_})
    sb = sample_box.to(_I32).contiguous() if sample_box is not None else None
    bs = box_sem.to(_I32).contiguous() if box_sem is not None else None
    bi = box_inst.to(_I32).contiguous() if box_inst is not None else None
    _capi.check(_capi.lib().pnr_composite(
        _capi.ptr(raw), _capi.ptr(z_vals), _capi.ptr(rays), R, N, Cn, Kn, int(bool(white_bkgd)),
        int(sem_activation == "softmax"), int(bool(mask_outside)), _capi.ptr(sb), _capi.ptr(bs), _capi.ptr(bi),
        B, C.byref(co), _capi.stream_ptr()), "pnr_composite")
    return out


@_on_tensor_device
def raw2outputs_backward(raw, z_vals, rays_d, grads: Dict[str, torch.Tensor], white_bkgd: bool = False,
                         num_classes: int = 0, num_instances: int = 0, sem_activation: str = "none",
                         sample_box: Optional[torch.Tensor] = None, box_sem: Optional[torch.Tensor] = None,
                         box_inst: Optional[torch.Tensor] = None, mask_outside: bool = False) -> torch.Tensor:
    """Gradient of a scalar loss with respect to `raw` [R,N,4+C+K], given its gradients with respect to the maps
    `raw2outputs` returns (`grads`: any subset of rgb_map, depth_map, acc_map, weights, semantic_map,
    instance_map, fixed_semantic_map, fixed_instance_map; `disp_map` is not differentiated).  First stage of
    the backward chain of the render path; checked against autograd through the oracle."""
    raw, z_vals = _f(raw, "raw"), _f(z_vals, "z_vals")
    R, N = z_vals.shape
    Cn, Kn = int(num_classes), int(num_instances)
    if raw.shape[-1] != 4 + Cn + Kn:
        raise ValueError(f"raw2outputs_backward: raw has {raw.shape[-1]} channels, expected {4 + Cn + Kn}")
    if "disp_map" in grads:
        raise ValueError("raw2outputs_backward: disp_map is not differentiated")
    rays = torch.cat([torch.zeros_like(rays_d), rays_d], -1) if rays_d.shape[-1] == 3 else rays_d
    rays = _f(rays, "rays")
    shapes = {"rgb_map": (R, 3), "depth_map": (R,), "acc_map": (R,), "weights": (R, N), "semantic_map": (R, Cn),
              "instance_map": (R, Kn), "fixed_semantic_map": (R, Cn), "fixed_instance_map": (R, Kn)}
    held = {}
    for k, g in grads.items():
        if k not in shapes:
            raise ValueError(f"raw2outputs_backward: unknown map {k!r}")
        if g is None:
            continue
        g = _f(g, k)
        if tuple(g.shape) != shapes[k]:
            raise ValueError(f"raw2outputs_backward: grad of {k} has shape {tuple(g.shape)}, expected {shapes[k]}")
        held[k] = g
    cg = _capi.PnrCompositeGrads(**{k: _capi.ptr(held[k]) if k in held else None
                                    for k, _ in _capi.PnrCompositeGrads._fields_})
    B = _box_table_size(raw, sample_box, box_sem, box_inst)
    sb = sample_box.to(_I32).contiguous() if sample_box is not None else None
    bs = box_sem.to(_I32).contiguous() if box_sem is not None else None
    bi = box_inst.to(_I32).contiguous() if box_inst is not None else None
    d_raw = torch.empty_like(raw)
    _capi.check(_capi.lib().pnr_composite_backward(
        _capi.ptr(raw), _capi.ptr(z_vals), _capi.ptr(rays), R, N, Cn, Kn, int(bool(white_bkgd)),
        int(sem_activation == "softmax"), int(bool(mask_outside)), _capi.ptr(sb), _capi.ptr(bs), _capi.ptr(bi),
        B, C.byref(cg), _capi.ptr(d_raw), _capi.stream_ptr()), "pnr_composite_backward")
    return d_raw


class _Raw2OutputsFn(torch.autograd.Function):
    """`raw2outputs` as an autograd node: losses written in torch on the composited maps back-propagate to `raw`
    through `pnr_composite_backward`.  The maps are returned in a fixed key order."""
    KEYS = ("rgb_map", "depth_map", "acc_map", "weights", "semantic_map", "instance_map",
            "fixed_semantic_map", "fixed_instance_map")

    @staticmethod
    def forward(ctx, raw, z_vals, rays_d, kw):
        out = raw2outputs(raw.detach(), z_vals, rays_d, **kw)
        ctx.save_for_backward(raw.detach(), z_vals, rays_d)
        ctx.kw = kw
        ctx.present = [k for k in _Raw2OutputsFn.KEYS if k in out]
        ctx.mark_non_differentiable(out["disp_map"])
        return tuple(out[k] for k in ctx.present) + (out["disp_map"],)

    @staticmethod
    def backward(ctx, *gs):
        raw, z_vals, rays_d = ctx.saved_tensors
        grads = {k: g for k, g in zip(ctx.present, gs[:-1]) if g is not None}
        return raw2outputs_backward(raw, z_vals, rays_d, grads, **ctx.kw), None, None, None


def raw2outputs_autograd(raw, z_vals, rays_d, white_bkgd: bool = False, num_classes: int = 0,
                         num_instances: int = 0, sem_activation: str = "none",
                         sample_box: Optional[torch.Tensor] = None, box_sem: Optional[torch.Tensor] = None,
                         box_inst: Optional[torch.Tensor] = None, mask_outside: bool = False) -> Dict[str, torch.Tensor]:
    """`raw2outputs` whose outputs carry gradients back to `raw` (z_vals and rays are treated as constants, as in
    the reference's training step where the sampler is not differentiated)."""
    kw = dict(white_bkgd=white_bkgd, num_classes=num_classes, num_instances=num_instances,
              sem_activation=sem_activation, sample_box=sample_box, box_sem=box_sem, box_inst=box_inst,
              mask_outside=mask_outside)
    res = _Raw2OutputsFn.apply(raw, z_vals, rays_d, kw)
    Cn, Kn = int(num_classes), int(num_instances)
    present = ["rgb_map", "depth_map", "acc_map", "weights"]
    if Cn > 0:
        present.append("semantic_map")
    if Kn > 0:
        present.append("instance_map")
    if sample_box is not None and box_sem is not None and Cn > 0:
        present.append("fixed_semantic_map")
    if sample_box is not None and box_inst is not None and Kn > 0:
        present.append("fixed_instance_map")
    out = dict(zip(present, res[:-1]))
    out["disp_map"] = res[-1]
    return out


@_on_tensor_device
def sample_pdf(z, weights, N_importance: int, det: bool = True, u: Optional[torch.Tensor] = None,
               want_idx: bool = False):
    """a10 on coarse depths z [R,N] and coarse weights [R,N] (bins = mid points, pdf = weights[1:-1]).
    Returns (z_fine [R,Ni], z_all [R,N+Ni] sorted[, idx [R,Ni] i64])."""
    z, weights = _f(z, "z"), _f(weights, "weights")
    R, N = z.shape
    Ni = int(N_importance)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, Ni).to(z.device)[None].expand(R, Ni)   # host linspace = oracle's
        else:
            u = torch.rand(R, Ni, device=z.device)
    u = _f(u, "u")
    z_f = torch.empty(R, Ni, dtype=_F32, device=z.device)
    z_all = torch.empty(R, N + Ni, dtype=_F32, device=z.device)
    idx = torch.empty(R, Ni, dtype=torch.int64, device=z.device) if want_idx else None
    _capi.check(_capi.lib().pnr_sample_pdf(_capi.ptr(z), _capi.ptr(weights), R, N, Ni, _capi.ptr(u),
                                           _capi.ptr(z_f), _capi.ptr(idx), _capi.ptr(z_all), _capi.stream_ptr()),
                "pnr_sample_pdf")
    return (z_f, z_all, idx) if want_idx else (z_f, z_all)


# ------------------------------------------------------------------------------------------------
# Renderer
# ------------------------------------------------------------------------------------------------
class Renderer:
    """Renderer(cfg, net).render(batch) -> dict of per-ray maps (a3).  batch keys: rays [R,6] (o||d),
    optional near/far [R], scene_aabb [2,3], box_center/box_half [B,3], box_rot [B,3,3], box_sem/box_inst [B],
    u [R,N] / u_fine [R,Ni] (externally supplied jitter), perturb."""

    def __init__(self, cfg, net, net_fine=None):
        self.cfg, self.net = cfg, net
        self.net_fine = net_fine if net_fine is not None else net
        self._t_vals = {}
        self._ws = {}                 # pnr_render_fused scratch, one buffer per device
        N, Ni = int(cfg.N_samples), int(getattr(cfg, "N_importance", 0))
        if N < 1 or N + Ni > MAX_SAMPLES_PER_RAY:
            raise ValueError(f"Renderer: N_samples + N_importance = {N} + {Ni} exceeds the {MAX_SAMPLES_PER_RAY} samples "
                             "per ray the compositing kernel handles")

    def _tv(self, N: int, device) -> torch.Tensor:
        key = (N, str(device))
        if key not in self._t_vals:
            self._t_vals[key] = torch.linspace(0.0, 1.0, N).to(device)   # CPU linspace, as the oracle's
        return self._t_vals[key]

    def _auto_chunk(self, rays) -> int:
        """Largest ray chunk whose intermediates (raw is the big one: N x (4+C+K) floats per ray, coarse + fine)
        fit in half of the free device memory; usually the whole frame (results do not depend on it)."""
        cfg = self.cfg
        N, Ni = int(cfg.N_samples), int(getattr(cfg, "N_importance", 0))
        ch = 4 + int(getattr(cfg, "num_classes", 0)) + int(getattr(cfg, "num_instances", 0))
        per_ray = 4 * ((N + (N + Ni if Ni else 0)) * (ch + 4) + 64)
        free, _ = torch.cuda.mem_get_info(rays.device)
        chunk = max(1024, int(0.5 * free // per_ray) // 1024 * 1024)
        return min(rays.shape[0], chunk)

    # -- a4: results are invariant to `chunk`; None renders every ray in one pass of persistent kernels
    def batchify_rays(self, rays, near, far, batch, chunk: Optional[int] = None):
        R = rays.shape[0]
        chunk = int(chunk) if chunk else R
        if chunk >= R:
            return self.render_rays(rays, near, far, batch, slice(0, R))
        outs = []
        for i in range(0, R, chunk):
            sl = slice(i, min(i + chunk, R))
            outs.append(self.render_rays(rays[sl].contiguous(), near[sl].contiguous(), far[sl].contiguous(),
                                         batch, sl))
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    def render_rays(self, rays, near, far, batch, sl):
        cfg = self.cfg
        N, Ni = int(cfg.N_samples), int(getattr(cfg, "N_importance", 0))
        Cn, Kn = int(getattr(cfg, "num_classes", 0)), int(getattr(cfg, "num_instances", 0))
        M = int(getattr(cfg, "max_hits", 4))
        perturb = float(batch.get("perturb", getattr(cfg, "perturb", 0.0)))
        out = {}
        has_boxes = "box_center" in batch and batch["box_center"].shape[0] > 0
        box_id = t_in = t_out = None
        if has_boxes:
            hit, box_id, t_in, t_out = intersect(rays, batch["box_center"], batch["box_half"],
                                                 batch["box_rot"], M)
            out.update(hit_mask=hit, box_id=box_id, t_in=t_in, t_out=t_out)
            if bool(getattr(cfg, "bound_by_primitives", False)):
                near, far = bound_by_primitives(hit, box_id, t_in, t_out, near, far)
        u = batch["u"][sl] if "u" in batch else None
        if has_boxes and str(getattr(cfg, "sample_mode", "uniform")) == "intervals":
            z, sb = interval_z(near, far, self._tv(N, rays.device), box_id, t_in, t_out, perturb, u)
        else:
            res = stratified_z(near, far, self._tv(N, rays.device), perturb, u, box_id, t_in, t_out,
                               want_tags=has_boxes)
            z, sb = res if has_boxes else (res, None)
        kw = dict(white_bkgd=bool(getattr(cfg, "white_bkgd", False)), num_classes=Cn, num_instances=Kn,
                  sem_activation=str(getattr(cfg, "sem_activation", "none")),
                  mask_outside=bool(getattr(cfg, "mask_outside", False)))
        if has_boxes:
            kw.update(box_sem=batch.get("box_sem"), box_inst=batch.get("box_inst"))
        raw = self.net.forward_rays(rays, z)
        res = raw2outputs(raw, z, rays, sample_box=sb, **kw)
        if Ni > 0:
            for k, v in res.items():
                out[k + "_0"] = v
            out["z_vals_0"] = z
            u_f = batch["u_fine"][sl] if "u_fine" in batch else None
            _, z = sample_pdf(z, res["weights"], Ni, det=(perturb == 0.0), u=u_f)
            sb = tag_samples(z, box_id, t_in, t_out) if has_boxes else None
            raw = self.net_fine.forward_rays(rays, z)
            res = raw2outputs(raw, z, rays, sample_box=sb, **kw)
        out.update(res)
        out["z_vals"] = z
        if sb is not None:
            out["sample_box"] = sb
        if bool(getattr(cfg, "return_raw", False)):
            out["raw"] = raw
        return out

    # -- a3
    @torch.no_grad()
    def render(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        if "rays" not in batch and "c2w" in batch:     # camera given instead of rays: generate them on the device
            dev = next(self.net.parameters()).device
            batch = dict(batch)
            batch["rays"] = generate_rays(int(cfg.H), int(cfg.W_img), batch.get("intrinsics", (cfg.fx, cfg.fy, cfg.cx, cfg.cy)),
                                          batch["c2w"], getattr(cfg, "camera", "pinhole"),
                                          int(batch.get("row0", 0)), batch.get("rows"), dev)
        rays = _f(batch["rays"], "batch['rays']")
        if "near" in batch and "far" in batch:
            near, far = _f(batch["near"], "near"), _f(batch["far"], "far")
        elif "scene_aabb" in batch:
            near, far = scene_near_far(rays, batch["scene_aabb"], float(cfg.near), float(cfg.far))
        else:
            near = torch.full((rays.shape[0],), float(cfg.near), dtype=_F32, device=rays.device)
            far = torch.full((rays.shape[0],), float(cfg.far), dtype=_F32, device=rays.device)
        staged = str(getattr(cfg, "render_path", "fused")) == "staged" or bool(getattr(cfg, "return_raw", False))
        with torch.cuda.device(rays.device):
            if staged:     # stage by stage from Python (one libpnr call per stage and chunk; keeps `raw`)
                chunk = getattr(cfg, "gpu_chunk", None) or self._auto_chunk(rays)
                out = self.batchify_rays(rays, near, far, batch, chunk)
                out["near"], out["far"] = near, far
            else:          # the whole frame in ONE libpnr call (pnr_render_fused), chunked inside by the workspace
                out = self.render_fused(rays, near, far, batch)
            if bool(getattr(cfg, "check_range", True)):
                self.net.check_range()
                if self.net_fine is not self.net:
                    self.net_fine.check_range()
        return out

    # -- a3/a4 through the single C-ABI entry point
    def _workspace(self, ctx, R: int, N: int, Ni: int, device) -> torch.Tensor:
        want = int(getattr(self.cfg, "workspace_mb", 0)) << 20
        chunk = int(getattr(self.cfg, "gpu_chunk", 0) or 0)
        if want <= 0 and chunk > 0:     # rays per chunk given instead of bytes
            one = int(_capi.lib().pnr_workspace_bytes(ctx, 1, N, Ni))
            want = int(_capi.lib().pnr_workspace_bytes(ctx, min(chunk, 1 << 16), N, Ni))
            want += max(0, chunk - (1 << 16)) * one
        if want <= 0:
            want = int(_capi.lib().pnr_workspace_bytes(ctx, R, N, Ni))
        ws = self._ws.get(str(device))
        if ws is None or ws.numel() < want:
            ws = torch.empty(want, dtype=torch.uint8, device=device)
            self._ws[str(device)] = ws
        return ws

    def render_fused(self, rays, near, far, batch) -> Dict[str, torch.Tensor]:
        cfg = self.cfg
        dev = rays.device
        R = rays.shape[0]
        N, Ni = int(cfg.N_samples), int(getattr(cfg, "N_importance", 0))
        Nt = N + Ni
        Cn, Kn = int(getattr(cfg, "num_classes", 0)), int(getattr(cfg, "num_instances", 0))
        M = int(getattr(cfg, "max_hits", 4))
        perturb = float(batch.get("perturb", getattr(cfg, "perturb", 0.0)))
        has_boxes = "box_center" in batch and batch["box_center"].shape[0] > 0
        ctx = self.net.pack(dev)
        ctx_fine = self.net_fine.pack(dev) if self.net_fine is not self.net else None
        e = lambda *shape, dtype=_F32: torch.empty(*shape, dtype=dtype, device=dev)

        def maps(n_samples):
            m = {"rgb_map": e(R, 3), "depth_map": e(R), "acc_map": e(R), "disp_map": e(R), "weights": e(R, n_samples)}
            if Cn > 0:
                m["semantic_map"] = e(R, Cn)
            if Kn > 0:
                m["instance_map"] = e(R, Kn)
            if has_boxes and batch.get("box_sem") is not None and Cn > 0:
                m["fixed_semantic_map"] = e(R, Cn)
            if has_boxes and batch.get("box_inst") is not None and Kn > 0:
                m["fixed_instance_map"] = e(R, Kn)
            return m

        def cstruct(m):
            return _capi.PnrCompositeOut(**{k: _capi.ptr(m[k]) if k in m else None
                                            for k, _ in _capi.PnrCompositeOut._fields_})
        final, coarse = maps(Nt), (maps(N) if Ni > 0 else {})
        keep = [rays, near, far]                      # tensors the call reads: alive until it is enqueued
        a = _capi.PnrRenderArgs()
        a.rays, a.R, a.near, a.far = _capi.ptr(rays), R, _capi.ptr(near), _capi.ptr(far)
        a.near_min, a.far_default = float(cfg.near), float(cfg.far)
        out: Dict[str, torch.Tensor] = {}
        if has_boxes:
            bc, bh, br = _f(batch["box_center"], "box_center"), _f(batch["box_half"], "box_half"), _f(batch["box_rot"], "box_rot")
            bs = batch["box_sem"].to(dev, _I32).contiguous() if batch.get("box_sem") is not None else None
            bi = batch["box_inst"].to(dev, _I32).contiguous() if batch.get("box_inst") is not None else None
            _box_table_size(rays, torch.empty(0, device=dev), bs, bi)
            keep += [bc, bh, br, bs, bi]
            a.box_center, a.box_half, a.box_rot = _capi.ptr(bc), _capi.ptr(bh), _capi.ptr(br)
            a.box_sem, a.box_inst, a.B, a.M = _capi.ptr(bs), _capi.ptr(bi), bc.shape[0], M
            hit8 = e(R, dtype=torch.uint8)
            out.update(box_id=e(R, M, dtype=_I32), t_in=e(R, M), t_out=e(R, M), sample_box=e(R, Nt, dtype=_I32))
            a.hit_mask, a.box_id, a.t_in, a.t_out = _capi.ptr(hit8), _capi.ptr(out["box_id"]), _capi.ptr(out["t_in"]), _capi.ptr(out["t_out"])
            a.sample_box = _capi.ptr(out["sample_box"])
        a.N, a.Ni = N, Ni
        tv = self._tv(N, dev)
        a.t_vals = _capi.ptr(tv)
        if perturb > 0.0:
            u = _f(batch["u"], "u") if "u" in batch else torch.rand(R, N, device=dev, dtype=_F32)
            keep.append(u)
            a.u = _capi.ptr(u)
        a.perturb = perturb
        if Ni > 0:
            if "u_fine" in batch:
                uf = _f(batch["u_fine"], "u_fine")
                a.u_fine_stride = Ni
            elif perturb == 0.0:
                uf = self._tv(Ni, dev)                 # deterministic sampler: one host-linspace row for every ray
                a.u_fine_stride = 0
            else:
                uf = torch.rand(R, Ni, device=dev, dtype=_F32)
                a.u_fine_stride = Ni
            keep.append(uf)
            a.u_fine = _capi.ptr(uf)
        a.sample_mode = _capi.SAMPLE_MODE[str(getattr(cfg, "sample_mode", "uniform"))] if has_boxes else 0
        a.white_bkgd = int(bool(getattr(cfg, "white_bkgd", False)))
        a.sem_softmax = int(str(getattr(cfg, "sem_activation", "none")) == "softmax")
        a.mask_outside = int(bool(getattr(cfg, "mask_outside", False)))
        a.bound_by_primitives = int(bool(getattr(cfg, "bound_by_primitives", False)))
        a.out, a.out0 = cstruct(final), cstruct(coarse)
        out["z_vals"] = e(R, Nt)
        a.z_vals = _capi.ptr(out["z_vals"])
        if Ni > 0:
            out["z_vals_0"] = e(R, N)
            a.z_vals0 = _capi.ptr(out["z_vals_0"])
        out["near"], out["far"] = near, far          # as given (bound_by_primitives works on a private copy)
        ws = self._workspace(ctx, R, N, Ni, dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        _capi.check(_capi.lib().pnr_render_fused(ctx, ctx_fine, C.byref(a), _capi.stream_ptr()), "pnr_render_fused")
        del keep
        if has_boxes:
            out["hit_mask"] = hit8.bool()
        out.update({k + "_0": v for k, v in coarse.items()})
        out.update(final)
        return out
