"""Where a training step's time goes: torch profiler over three steps (python tools/profile_train_step.py [preset])."""
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
R, N = 2048, 192
cfg = PN.make_cfg(preset)
net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
g = torch.Generator().manual_seed(0)
rays = torch.cat([torch.randn(R, 3, generator=g) * 0.5, torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)], -1).to(DEV)
z = torch.sort(torch.rand(R, N, generator=g) * 6 + 0.5, -1).values.to(DEV)
batch = {"rgb": torch.rand(R, 3, generator=g).to(DEV), "depth": (torch.rand(R, generator=g) * 6).to(DEV)}
if cfg.num_classes:
    batch["pseudo_label"] = torch.randint(-1, cfg.num_classes, (R,), generator=g).to(DEV)
opt = torch.optim.Adam(net.parameters(), lr=5e-4)
for _ in range(3):
    opt.zero_grad(set_to_none=True); training_step(net, rays, z, batch, (1.0, 0.1, 1.0, 0.0)); opt.step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        opt.zero_grad(set_to_none=True); training_step(net, rays, z, batch, (1.0, 0.1, 1.0, 0.0)); opt.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
