"""Backward of `Network.forward`: dL/d(parameters) from dL/draw.

The work is split where the network splits:
  * the trunk (`pts_linears`: D of the network's GEMM layers, ~80 % of its FLOPs) runs on the fused tensor-core kernel:
    `pnr_mlp_trunk_forward` gives the trunk output h, `pnr_mlp_backward_trunk` recomputes the trunk per tile, runs
    the layers in reverse with the transposed weight stream and keeps every operand of the weight-gradient GEMMs
    (activations H_i, pre-activation gradients dZ_j) in one fp32 stash;
  * the layers after the trunk (alpha / feature / view / rgb / the two heads: small GEMMs with K <= W) are
    differentiated by torch on h - plain library GEMMs, which is also what the trunk's dW_j = dZ_j^T H_{j-1} are.
Nothing here imports the oracle; tests compare every parameter's gradient with autograd through the oracle network."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from ..networks.renderer import panopticnerf_renderer as P


def _tf32_parts(x: torch.Tensor):
    """x = hi + lo with hi exactly representable in TF32 (10 mantissa bits; the low 13 bits cleared)."""
    hi = (x.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, x - hi


def matmul_3xtf32(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
    """op(a) @ op(b) to ~2^-20 per product on the TF32 tensor cores: hi.hi + lo.hi + hi.lo with fp32 accumulation (the
    same 3-pass split the fused kernel uses with fp16 parts, here with TF32's fp32 exponent range: no scaling needed).
    The library's fp32 GEMM without tensor cores is ~10x slower and would dominate a training step.  The transposes
    are views (the BLAS takes them as operand flags): a [S, W] gradient is never copied into [W, S]."""
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        ah, al = _tf32_parts(a)
        bh, bl = _tf32_parts(b)
        if trans_a:
            ah, al = ah.t(), al.t()
        if trans_b:
            bh, bl = bh.t(), bl.t()
        return ah @ bh + (al @ bh + ah @ bl)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


class _Linear3x(torch.autograd.Function):
    """F.linear whose three GEMMs (forward, dL/dx, dL/dW) are 3xTF32: the layers after the trunk at tensor-core speed
    with fp32-grade results."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return matmul_3xtf32(x, weight, trans_b=True) + bias

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = g.contiguous()
        return matmul_3xtf32(g, weight), matmul_3xtf32(g, x, trans_a=True), g.sum(0)


def _lin(layer, x):
    return _Linear3x.apply(x, layer.weight, layer.bias)


def _tail(net, h: torch.Tensor, ed: torch.Tensor) -> torch.Tensor:
    """raw from the trunk output: the reference Network.forward after `pts_linears` (same module names)."""
    sigma = _lin(net.alpha_linear, h)
    feat = _lin(net.feature_linear, h)
    g = F.relu(_lin(net.views_linears[0], torch.cat([feat, ed], -1)))
    outs = [_lin(net.rgb_linear, g), sigma]
    if net.C > 0:
        outs.append(_lin(net.semantic_linears[1], F.relu(_lin(net.semantic_linears[0], h))))
    if net.K > 0:
        outs.append(_lin(net.instance_linears[1], F.relu(_lin(net.instance_linears[0], h))))
    return torch.cat(outs, -1)


def network_backward(net, d_raw: torch.Tensor, pts: Optional[torch.Tensor] = None, viewdirs: Optional[torch.Tensor] = None,
                     rays: Optional[torch.Tensor] = None, z: Optional[torch.Tensor] = None,
                     return_input_grad: bool = False) -> Dict[str, torch.Tensor]:
    """{parameter name: gradient} of `net` for dL/draw = d_raw [S, 4+C+K], samples given as (pts, viewdirs) [S,3] each
    or as (rays [R,6], z [R,N]).  With return_input_grad also 'embedded_xyz': dL/d gamma(x) [S, 3+6*xyz_res]."""
    if pts is None:
        o, d = rays[:, None, :3], rays[:, None, 3:]
        pts_ = (o + d * z[..., None]).reshape(-1, 3)                 # the kernel forms the same points (mul, then add)
        vd = (rays[:, 3:] / rays[:, 3:].norm(dim=-1, keepdim=True))[:, None, :].expand(-1, z.shape[1], -1).reshape(-1, 3)
    else:
        pts_, vd = pts.reshape(-1, 3), viewdirs.reshape(-1, 3)
    S_ = pts_.shape[0]
    d_raw = d_raw.reshape(S_, -1).to(torch.float32)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False                    # the library GEMMs below are fp32
    try:
        h = net.trunk_forward(pts=pts, rays=rays, z=z).requires_grad_(True)
        ed = P.embed(vd.contiguous(), net.Ld)
        tail_named = [(n, p) for n, p in net.named_parameters() if not n.startswith("pts_linears.")]
        with torch.enable_grad():
            raw = _tail(net, h, ed)
            g = torch.autograd.grad(raw, [h] + [p for _, p in tail_named], d_raw, allow_unused=True)
        grads = {n: (gi if gi is not None else torch.zeros_like(p)) for (n, p), gi in zip(tail_named, g[1:])}
        d_emb, st = net.backward_trunk(g[0].contiguous(), pts=pts, rays=rays, z=z, stash=True)
        ex = P.embed(pts_.contiguous(), net.Lx)
        D = net.D
        for j in range(D):
            dZ = st[2 * D - 2 - j]
            inp = ex if j == 0 else (torch.cat([ex, st[j - 1]], -1) if j == net.skip + 1 else st[j - 1])
            grads[f"pts_linears.{j}.weight"] = matmul_3xtf32(dZ, inp, trans_a=True)
            grads[f"pts_linears.{j}.bias"] = dZ.sum(0)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    if return_input_grad:
        grads["embedded_xyz"] = d_emb
    return grads


class _NetworkFn(torch.autograd.Function):
    """raw = Network.forward(pts, viewdirs) on the fused kernel, differentiable w.r.t. the network's parameters."""

    @staticmethod
    def forward(ctx, net, pts, viewdirs, *params):
        ctx.net, ctx.names = net, [n for n, _ in net.named_parameters()]
        ctx.save_for_backward(pts, viewdirs)
        with torch.no_grad():
            return net.forward(pts, viewdirs)

    @staticmethod
    def backward(ctx, d_raw):
        pts, viewdirs = ctx.saved_tensors
        g = network_backward(ctx.net, d_raw.contiguous(), pts=pts.reshape(-1, 3).contiguous(),
                             viewdirs=viewdirs.reshape(-1, 3).contiguous())
        return (None, None, None) + tuple(g[n] for n in ctx.names)


class _NetworkRaysFn(torch.autograd.Function):
    """raw [R,N,CH] = Network.forward_rays(rays, z) on the fused kernel, differentiable w.r.t. the parameters."""

    @staticmethod
    def forward(ctx, net, rays, z, *params):
        ctx.net, ctx.names = net, [n for n, _ in net.named_parameters()]
        ctx.save_for_backward(rays, z)
        with torch.no_grad():
            return net.forward_rays(rays, z)

    @staticmethod
    def backward(ctx, d_raw):
        rays, z = ctx.saved_tensors
        g = network_backward(ctx.net, d_raw.contiguous(), rays=rays, z=z)
        return (None, None, None) + tuple(g[n] for n in ctx.names)


def network_forward_rays_autograd(net, rays: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """`net.forward_rays(rays, z)` (points and view directions formed in the kernel) with gradients to the parameters."""
    return _NetworkRaysFn.apply(net, rays.contiguous(), z.contiguous(), *[p for _, p in net.named_parameters()])


def training_step(net, rays: torch.Tensor, z: torch.Tensor, batch: Dict[str, torch.Tensor], weights=(1.0, 0.1, 1.0, 1.0),
                  white_bkgd: bool = False, sem_is_prob: bool = False, sample_box: Optional[torch.Tensor] = None,
                  box_sem: Optional[torch.Tensor] = None, mask_outside: bool = False):
    """One pass of the reference NetworkWrapper's job for one set of samples: raw = net(rays, z) (fused MLP kernel),
    maps = raw2outputs (compositing kernel), loss terms (loss kernel), and back: dL/dmaps (loss kernel) -> dL/draw
    (pnr_composite_backward) -> dL/dparameters (network_backward).  Leaves the gradients in `p.grad` like
    `loss.backward()` does; returns (total, terms).  Depths z are treated as constants (the sampler is not
    differentiated, as in the reference)."""
    from .losses import panoptic_losses
    raw = network_forward_rays_autograd(net, rays, z)
    out = P.raw2outputs_autograd(raw, z, rays[:, 3:].contiguous(), white_bkgd=white_bkgd, num_classes=net.C,
                                 num_instances=net.K, sample_box=sample_box, box_sem=box_sem, mask_outside=mask_outside)
    total, terms = panoptic_losses(out, batch, weights, sem_is_prob=sem_is_prob)
    total.backward()
    return total.detach(), terms


def network_forward_autograd(net, pts: torch.Tensor, viewdirs: torch.Tensor) -> torch.Tensor:
    """`net(pts, viewdirs)` whose result back-propagates into `net.parameters()` through `network_backward`."""
    return _NetworkFn.apply(net, pts, viewdirs, *[p for _, p in net.named_parameters()])
