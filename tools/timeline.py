"""Dump the clock64 timeline of one tile of the fused MLP kernel (block 0, third tile).
Needs the -DPNR_TIMELINE build: python -m panopticnerf_b200._build --timeline ; PNR_LIB=panopticnerf_b200/libpnr_timeline.so python tools/timeline.py"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import panopticnerf_b200 as PN
from panopticnerf_b200 import _capi, synthetic as S
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P
backward = "--backward" in sys.argv          # the trunk-backward kernel instead of the forward one
composite = "--composite" in sys.argv        # the forward kernel with the compositing epilogue
sys.argv = [a for a in sys.argv if a not in ("--backward", "--composite")]
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
preset = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
cfg = PN.make_cfg(preset, precision=prec)
dev = "cuda:0"
net = S.init_network_weights(PN.make_network(cfg)).to(dev)
batch = {k: v.to(dev) for k, v in S.make_batch(cfg, rows=40).items()}
rays = batch["rays"]
near, far = P.scene_near_far(rays, batch["scene_aabb"], cfg.near, cfg.far)
z = P.stratified_z(near, far, torch.linspace(0, 1, cfg.N_samples).to(dev))
ctx = net.pack(dev)
raw = torch.empty(rays.shape[0], cfg.N_samples, net.out_channels, device=dev)
tl = torch.zeros(8192, dtype=torch.int64, device=dev)
if backward:
    grad_h = torch.randn(z.numel(), cfg.W, device=dev)
    _capi.check(_capi.lib().pnr_debug_timeline(ctx, tl.data_ptr()))
    for _ in range(2):
        net.backward_trunk(grad_h, rays=rays, z=z)
    _capi.check(_capi.lib().pnr_debug_timeline(ctx, None))
if composite:
    _capi.check(_capi.lib().pnr_debug_timeline(ctx, tl.data_ptr()))
    for _ in range(2):
        net.forward_composite(rays, z)
    _capi.check(_capi.lib().pnr_debug_timeline(ctx, None))
for _ in range(0 if (backward or composite) else 2):
    _capi.check(_capi.lib().pnr_mlp_forward_timeline(ctx, rays.data_ptr(), z.data_ptr(), rays.shape[0], cfg.N_samples,
                                                     raw.data_ptr(), tl.data_ptr(), _capi.stream_ptr()))
torch.cuda.synchronize()
t = tl.cpu().tolist()
mma = [t[i * 5:(i + 1) * 5] for i in range(800) if t[i * 5] > 0]
t0 = mma[0][0]
print(f"precision {prec}: {len(mma)} stages; tile span {mma[-1][4] - t0} cycles")
print("stage: arrive  ready  issued   (waits=ready-arrive, issue=issued-ready) | tma_issue")
for i, (a, r, m, c, d) in enumerate(mma):
    print(f"{i:4d} {a - t0:8d}  wait {r - a:5d} mma {m - r:5d} commit {c - m:5d} sync {d - c:5d} | next-gap {(mma[i + 1][0] - d) if i + 1 < len(mma) else 0:5d} | tma {t[6144 + i] - t0:8d}")
print("step half: wait_start acc_ready done  (wait, work)")
for k in range(2 * 24):
    w, a, d = t[4096 + k * 3:4096 + k * 3 + 3]
    if w > 0:
        print(f"{k // 2:3d} h{k % 2} {w - t0:8d} {a - t0:8d} {d - t0:8d}   wait {a - w:6d} work {d - a:6d}")

