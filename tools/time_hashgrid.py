"""Time the hash-grid encoder and the panoptic fusion kernel at frame size:
python tools/time_hashgrid.py [n_points_millions]   (default 33.9 = the samples of one cfg2 frame)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from panopticnerf_b200.lib.networks.encoding import HashGrid
from panopticnerf_b200.lib.visualizers import fuse_panoptic

DEV = "cuda:0"
n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 33_882_112


def timed(fn, reps=5):
    ts = []
    for i in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for L, F, T in ((16, 2, 19), (16, 2, 22)):
    enc = HashGrid(L, F, T, 16.0, 1.3819).to(DEV)
    # points along rays (coherent, like render samples) and uniformly random ones
    t = torch.linspace(0, 1, 64, device=DEV)
    o = torch.rand(n // 64, 1, 3, device=DEV) * 0.2 + 0.4
    d = torch.nn.functional.normalize(torch.randn(n // 64, 1, 3, device=DEV), dim=-1) * 0.4
    for name, x in (("ray samples", (o + d * t[None, :, None]).reshape(-1, 3).clamp(0, 1).contiguous()),
                    ("random", torch.rand(n // 64 * 64, 3, device=DEV))):
        ms = timed(lambda: enc(x))
        m = x.shape[0]
        print(f"hashgrid L={L} F={F} T=2^{T} ({enc.table.numel() * 4 / 2**20:.0f} MB table), {m / 1e6:.1f} M {name}: {ms:7.3f} ms  "
              f"{m / ms / 1e3:7.1f} M points/s  in+out {m * (12 + L * F * 4) / ms / 1e6:6.0f} GB/s  gathers {m * L * 8 * F * 4 / ms / 1e6:6.0f} GB/s", flush=True)
        del x
    del enc
R, C, K = 1408 * 376, 45, 64
out = {"semantic_map": torch.rand(R, C, device=DEV), "instance_map": torch.rand(R, K, device=DEV)}
thing = (torch.arange(C) % 2).to(torch.uint8)
ic = torch.arange(K) % C
pal = torch.randint(0, 256, (C, 3), dtype=torch.uint8)
ms = timed(lambda: fuse_panoptic(out, thing, ic, None, None, pal), reps=9)
print(f"panoptic_fuse {R} rays C={C} K={K}: {ms:.3f} ms  {R * (C + K) * 4 / ms / 1e6:.0f} GB/s read")
