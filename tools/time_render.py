"""Renderer.render time for BASELINE presets (CUDA events, L2 flushed between runs):
    python tools/time_render.py [preset[:rows] ...]      e.g.  cfg2 cfg3   or   cfg3:32  (a 32-row strip)"""
import json, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import panopticnerf_b200 as PN
from panopticnerf_b200 import synthetic as S
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for spec in sys.argv[1:] or ["cfg2", "cfg3"]:
    preset, _, rows = spec.partition(":")
    cfg = PN.make_cfg(preset)
    net = S.init_network_weights(PN.make_network(cfg)).to(dev)
    batch = {k: v.to(dev) for k, v in S.make_batch(cfg, rows=int(rows) if rows else None).items()}
    ren = PN.make_renderer(cfg, net)
    ts = []
    for i in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = ren.render(batch); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    R = batch["rays"].shape[0]
    ms = sum(ts) / len(ts)
    res[spec] = dict(ms=ms, rays_per_s=R / ms * 1e3, peak_mem_gb=torch.cuda.max_memory_allocated() / 2**30)
    print(spec, res[spec], flush=True)
    del out, batch, net, ren
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "render_times.json").write_text(json.dumps(res, indent=1))
