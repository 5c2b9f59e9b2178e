"""Time one training step (forward render of a ray batch + losses + backward to every parameter) on the library's
kernels: python tools/time_train_step.py [preset] [n_rays] [n_samples].  Default: cfg3 network, 2048 rays x 192."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import panopticnerf_b200 as PN
from panopticnerf_b200 import synthetic as S
from panopticnerf_b200.lib.train import training_step

DEV = "cuda:0"
preset = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
N = int(sys.argv[3]) if len(sys.argv) > 3 else 192
cfg = PN.make_cfg(preset)
net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
g = torch.Generator().manual_seed(0)
rays = torch.cat([torch.randn(R, 3, generator=g) * 0.5, torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)], -1).to(DEV)
z = torch.sort(torch.rand(R, N, generator=g) * 6 + 0.5, -1).values.to(DEV)
batch = {"rgb": torch.rand(R, 3, generator=g).to(DEV), "depth": (torch.rand(R, generator=g) * 6).to(DEV)}
if cfg.num_classes:
    batch["pseudo_label"] = torch.randint(-1, cfg.num_classes, (R,), generator=g).to(DEV)
opt = torch.optim.Adam(net.parameters(), lr=5e-4)
ts, parts = [], {}


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for i in range(8):
    opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    a = ev()
    total, terms = training_step(net, rays, z, batch, (1.0, 0.1, 1.0, 0.0))
    b = ev()
    opt.step()
    c = ev()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append((a.elapsed_time(b), b.elapsed_time(c)))
fb = sorted(t[0] for t in ts)[len(ts) // 2]
st = sorted(t[1] for t in ts)[len(ts) // 2]
flops = 3 * PN.mlp_flops_per_sample(cfg) * R * N if hasattr(PN, "mlp_flops_per_sample") else 0
print(f"{preset} train step, {R} rays x {N} samples ({R * N / 1e3:.0f} k samples): forward+loss+backward {fb:.2f} ms, optimizer {st:.2f} ms "
      f"-> {R / (fb + st) * 1e3 / 1e3:.1f} k rays/s, {R * N / (fb + st) / 1e3:.2f} M samples/s; loss {float(total):.4f}")
