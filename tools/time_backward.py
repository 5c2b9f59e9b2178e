"""Time the trunk-backward kernel (pnr_mlp_backward_trunk) on a strip of a cfg frame:
python tools/time_backward.py [preset] [rows] [precision ...].  The incoming gradient is [S, W] fp32 (1 KB per
sample at W = 256), so the strip is sized to keep it at a few GB."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import panopticnerf_b200 as PN
from panopticnerf_b200 import synthetic as S
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P

DEV = "cuda:0"
preset = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 64
precs = sys.argv[3:] or ["fp16x3"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for prec in precs:
    cfg = PN.make_cfg(preset, precision=prec)
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    batch = {k: v.to(DEV) for k, v in S.make_batch(cfg, rows=rows).items()}
    rays = batch["rays"]
    near, far = P.scene_near_far(rays, batch["scene_aabb"], cfg.near, cfg.far)
    z = P.stratified_z(near, far, torch.linspace(0, 1, cfg.N_samples).to(DEV))
    S_ = z.numel()
    grad_h = torch.randn(S_, cfg.W, device=DEV)
    ts, tf = [], []
    for i in range(7):
        for which, acc in (("b", ts), ("f", tf)):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            if which == "b":
                net.backward_trunk(grad_h, rays=rays, z=z, grad_scale=64.0)
            else:
                net.forward_rays(rays, z)
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                acc.append(a.elapsed_time(b))
    ts.sort(); tf.sort()
    D, W, Ex = cfg.D, cfg.W, 3 + 6 * cfg.xyz_res
    fl_f = 2 * (Ex * W + (D - 2) * W * W + (W + Ex) * W)          # trunk forward (recomputed)
    fl_b = 2 * ((D - 2) * W * W + W * (W + Ex) + W * Ex)          # trunk backward (data gradient)
    t = ts[len(ts) // 2]
    print(f"{preset} {rows} rows ({S_ / 1e6:.2f} M samples) {prec}: backward_trunk median {t:8.3f} ms "
          f"= {S_ * (fl_f + fl_b) / t / 1e9:7.1f} TFLOP/s algorithmic (fwd recompute + dX), "
          f"grad_h stream {S_ * W * 4 / t / 1e6:6.1f} GB/s; full forward of the same samples {tf[len(tf) // 2]:8.3f} ms", flush=True)
