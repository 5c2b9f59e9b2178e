"""Time only the fused MLP kernel on a cfg frame: python tools/time_mlp.py [preset] [precision ...]."""
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
precs = sys.argv[2:] or ["fp16x3"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for prec in precs:
    cfg = PN.make_cfg(preset, precision=prec)
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    batch = {k: v.to(DEV) for k, v in S.make_batch(cfg).items()}
    rays = batch["rays"]
    near, far = P.scene_near_far(rays, batch["scene_aabb"], cfg.near, cfg.far)
    z = P.stratified_z(near, far, torch.linspace(0, 1, cfg.N_samples).to(DEV))
    ts = []
    for i in range(9):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        net.forward_rays(rays, z)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{preset} mlp {prec:7s}: median {ts[len(ts) // 2]:8.3f} ms  best {ts[0]:8.3f} ms  "
          f"{rays.shape[0] / ts[len(ts) // 2] / 1e3:6.2f} Mrays/s", flush=True)
