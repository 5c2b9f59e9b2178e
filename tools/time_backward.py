"""Time pnr_composite_backward at the cfg2 / cfg3 frame sizes (HBM-bound: reads raw + z, writes d_raw)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P

DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, R, N, C, K in (("cfg2 rgb+sigma", 376 * 1408, 64, 0, 0), ("cfg3 heads, 100k rays", 100_000, 192, 45, 50)):
    raw = torch.randn(R, N, 4 + C + K, device=DEV)
    z = torch.sort(torch.rand(R, N, device=DEV) * 60 + 2, -1).values
    rays = torch.randn(R, 6, device=DEV)
    grads = {"rgb_map": torch.randn(R, 3, device=DEV), "depth_map": torch.randn(R, device=DEV)}
    if C:
        grads["semantic_map"] = torch.randn(R, C, device=DEV)
        grads["instance_map"] = torch.randn(R, K, device=DEV)
    ts = []
    for i in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        P.raw2outputs_backward(raw, z, rays, grads, num_classes=C, num_instances=K)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    ts.sort()
    byt = R * N * (8 * (4 + C + K) + 8) + R * 40
    print(f"composite_backward {name:24s}: {ts[len(ts) // 2]:7.3f} ms  {byt / ts[len(ts) // 2] / 1e6:8.1f} GB/s algorithmic "
          f"({byt / 1e6:.0f} MB)", flush=True)
