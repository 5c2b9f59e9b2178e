"""Collect the bench lines of tools/r2_scaling.sh (gpurun_out/scale_r2_<arm>_n<N>.json) into profiles/r02_scaling.md
and copy them to profiles/r02_scale_<arm>_n<N>.json."""
import json
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ARMS = [("weak_cfg2_maps", "cfg2, one frame per GPU (weak), fp32 map tiles (20 B/ray)"),
        ("strong_cfg2_maps", "cfg2, ONE frame ray-sharded (strong), fp32 map tiles"),
        ("strong_cfg2_labels", "cfg2, ONE frame ray-sharded (strong), label tiles (7 B/ray: rgb8 + depth)"),
        ("strong_cfg3_labels", "cfg3 (heads, coarse + fine), ONE frame ray-sharded, label tiles (11 B/ray)"),
        ("strong_cfg5_labels", "cfg5 equirect 2048x1024x192 with heads, ONE panorama ray-sharded (128 rows per GPU at N = 8), label tiles")]
rows = {}
for arm, _ in ARMS:
    for n in (1, 2, 4, 8):
        f = ROOT / "gpurun_out" / f"scale_r2_{arm}_n{n}.json"
        if not f.exists():
            continue
        try:
            d = json.loads(f.read_text().strip().splitlines()[-1])
        except Exception:
            continue
        rows[(arm, n)] = d
        shutil.copy(f, ROOT / "profiles" / f"r02_scale_{arm}_n{n}.json")
out = ["# Multi-GPU scaling, round 2 (one box, N x B200, NCCL over NVLink through `pnr_allgather_outputs`)", "",
       "`bench.py --gpus N --steps 6 --warmup 3 [--config ...] [--scaling strong] [--gather labels]`; device time, max over",
       "ranks, gather inside the step; `e2e` adds the H2D of the rays and the D2H of the gathered tiles.  Raw lines: `r02_scale_*.json`.", ""]
for arm, title in ARMS:
    have = [n for n in (1, 2, 4, 8) if (arm, n) in rows]
    if not have:
        continue
    out += [f"## {title}", "", "| N | rays/s | ms/step | x of N=1 | e2e rays/s | MLP kernel ms (per rank) | gathered bytes per rank | SM MHz |", "|---|---|---|---|---|---|---|---|"]
    base = rows.get((arm, 1)) or rows.get(("strong_cfg2_maps", 1) if "cfg2" in arm else (arm, have[0]))
    for n in have:
        d = rows[(arm, n)]
        g = d["config"]["gather"]
        gb = g.split(",")[-1].strip() if "bytes" in g else "-"
        rel = d["value"] / base["value"] if base else float("nan")
        out.append(f"| {n} | {d['value'] / 1e6:.2f} M | {d['ms_per_step']:.2f} | {rel:.2f} | {d['e2e']['value'] / 1e6:.2f} M | "
                   f"{d['roofline']['kernel_ms']:.2f} | {gb} | {d['clocks']['sm_mhz']} |")
    out.append("")
(ROOT / "profiles" / "r02_scaling.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
