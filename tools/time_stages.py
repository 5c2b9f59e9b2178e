"""Per-stage device timing of one frame (CUDA events on the launching stream, L2 flushed between
iterations).  Development tool: prints a table and writes gpurun_out/stage_times.json."""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import panopticnerf_b200 as PN  # noqa: E402
from panopticnerf_b200 import synthetic as S  # noqa: E402
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P  # noqa: E402

DEV = "cuda:0"


def flops_per_sample(cfg):
    D, W = cfg.D, cfg.W
    Ex, Ed = 3 + 6 * cfg.xyz_res, 3 + 6 * cfg.view_res
    mac = Ex * W + (D - 2) * W * W + (W + Ex) * W + W + W * W + (W + Ed) * (W // 2) + (W // 2) * 3
    if cfg.num_classes:
        mac += W * (W // 2) + (W // 2) * cfg.num_classes
    if cfg.num_instances:
        mac += W * (W // 2) + (W // 2) * cfg.num_instances
    return 2 * mac


def timeit(fn, iters=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else None
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for prec in ("fp16x3", "bf16x3", "fp16", "bf16"):
        cfg = PN.make_cfg(preset, precision=prec)
        net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
        batch = {k: v.to(DEV) for k, v in S.make_batch(cfg, rows=rows).items()}
        rays = batch["rays"]
        R, N = rays.shape[0], cfg.N_samples
        near, far = P.scene_near_far(rays, batch["scene_aabb"], cfg.near, cfg.far)
        t_vals = torch.linspace(0, 1, N).to(DEV)
        z = P.stratified_z(near, far, t_vals)
        fl = flops_per_sample(cfg) * R * N
        med, best = timeit(lambda: net.forward_rays(rays, z), flush=flush)
        out[f"mlp_{prec}"] = dict(ms=med, best_ms=best, tflops=fl / med / 1e9, rays_per_s=R / med * 1e3)
        print(f"{preset} mlp {prec:7s}: {med:8.3f} ms (best {best:.3f})  {fl / med / 1e9:8.1f} TFLOP/s alg  "
              f"{R / med * 1e3 / 1e6:6.2f} Mrays/s", flush=True)
        if prec == "fp16x3":
            raw = net.forward_rays(rays, z)
            hit = P.intersect(rays, batch["box_center"], batch["box_half"], batch["box_rot"], 4)
            grads = {"rgb_map": torch.randn(R, 3, device=DEV), "depth_map": torch.randn(R, device=DEV)}
            for name, fn, byt in [
                ("near_far", lambda: P.scene_near_far(rays, batch["scene_aabb"], cfg.near, cfg.far), R * 32),
                ("intersect", lambda: P.intersect(rays, batch["box_center"], batch["box_half"], batch["box_rot"], 4), R * (24 + 1 + 48)),
                ("stratified+tag", lambda: P.stratified_z(near, far, t_vals, 0.0, None, hit[1], hit[2], hit[3], want_tags=True), R * (8 + 48 + 8 * N)),
                ("composite", lambda: P.raw2outputs(raw, z, rays, num_classes=cfg.num_classes, num_instances=cfg.num_instances),
                 R * N * (4 * raw.shape[-1] + 8) + R * 20),
                ("composite_bwd", lambda: P.raw2outputs_backward(raw, z, rays, grads, num_classes=cfg.num_classes, num_instances=cfg.num_instances),
                 R * N * (8 * raw.shape[-1] + 8) + R * 20),
                ("render_total", lambda: PN.make_renderer(cfg, net).render(batch), 0),
            ]:
                med, best = timeit(fn, flush=flush)
                out[name] = dict(ms=med, best_ms=best, gbs=byt / med / 1e6 if byt else None)
                print(f"{preset} {name:15s}: {med:8.3f} ms (best {best:.3f})" + (f"  {byt / med / 1e6:8.1f} GB/s alg" if byt else ""), flush=True)
            x = torch.rand(4_000_000, 3, device=DEV) * 60
            med, best = timeit(lambda: P.embed(x, 10), flush=flush)
            byt = x.shape[0] * (12 + 63 * 4)
            out["encode"] = dict(ms=med, gbs=byt / med / 1e6)
            print(f"encode 4M x L=10   : {med:8.3f} ms  {byt / med / 1e6:8.1f} GB/s alg", flush=True)
            del x, raw
    os.makedirs(ROOT / "gpurun_out", exist_ok=True)
    (ROOT / "gpurun_out" / f"stage_times_{preset}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
