"""GPU: device ray generation, camera-driven render, and the full-size configurations
of BASELINE.json exercised through size-independent properties (the oracle cannot run them in seconds):
chunk invariance, weights in [0,1] summing to acc, sorted fine depths, shard == whole."""
import math

import pytest
import torch

import panopticnerf_b200 as PN
from oracle import reference_renderer as O
from panopticnerf_b200 import parallel, synthetic as S
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pose(yaw=0.3, t=(1.0, -2.0, 0.5)):
    c, s = math.cos(yaw), math.sin(yaw)
    return torch.tensor([[c, 0.0, s, t[0]], [0.0, 1.0, 0.0, t[1]], [-s, 0.0, c, t[2]]])


def test_generate_rays_pinhole_bit_exact_and_equirect():
    cfg = PN.make_cfg("cfg2")
    k = (cfg.fx, cfg.fy, cfg.cx, cfg.cy)
    ref = O.generate_rays(cfg.H, cfg.W_img, k, _pose(), "pinhole", row0=100, rows=7)
    got = P.generate_rays(cfg.H, cfg.W_img, k, _pose(), "pinhole", row0=100, rows=7, device=DEV)
    assert torch.equal(got.cpu(), ref)
    cfg5 = PN.make_cfg("cfg5")
    ref = O.generate_rays(cfg5.H, cfg5.W_img, k, _pose(), "equirect", row0=500, rows=3)
    got = P.generate_rays(cfg5.H, cfg5.W_img, k, _pose(), "equirect", row0=500, rows=3, device=DEV)
    assert torch.allclose(got.cpu(), ref, atol=2e-6)
    assert torch.allclose(got[:, 3:].norm(dim=-1).cpu(), torch.ones(got.shape[0]), atol=1e-5)
    assert P.generate_rays(cfg.H, cfg.W_img, k, _pose(), rows=0, device=DEV).shape == (0, 6)


def test_generate_rays_fisheye_bit_exact():
    """KITTI-360 fisheye (MEI) camera: every operation of the unprojection is a separately rounded fp32 op in a fixed
    order (8 Newton steps), so the device rays equal the oracle's bit for bit - also beyond the field of view."""
    k = (1336.3220825849971, 1335.7883350012958, 716.94323510126321, 705.76498308221585, 2.2134047507854890,
         1.6798235660113681e-02, 1.6548773243373522)
    ref = O.generate_rays(1400, 1400, k, _pose(), "fisheye", row0=0, rows=1400)
    got = P.generate_rays(1400, 1400, k, _pose(), "fisheye", row0=0, rows=1400, device=DEV)
    assert torch.equal(got.cpu(), ref)
    sub = P.generate_rays(1400, 1400, k, _pose(), "fisheye", row0=650, rows=9, device=DEV)
    assert torch.equal(sub.cpu(), ref[650 * 1400:659 * 1400])
    with pytest.raises(ValueError, match="7 intrinsics"):
        P.generate_rays(1400, 1400, k[:4], _pose(), "fisheye", device=DEV)


def test_render_from_camera_equals_render_from_rays():
    cfg = PN.make_cfg("cfg1")
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    ren = PN.make_renderer(cfg, net)
    pose = _pose(0.05, (0.0, 0.0, 0.0))
    rays = O.generate_rays(cfg.H, cfg.W_img, (cfg.fx, cfg.fy, cfg.cx, cfg.cy), pose)
    scene = {k: v.to(DEV) for k, v in S.make_batch(cfg).items() if k != "rays"}
    a = ren.render(dict(scene, rays=rays.to(DEV)))
    b = ren.render(dict(scene, c2w=pose))
    for key in ("rgb_map", "depth_map", "acc_map", "z_vals", "hit_mask"):
        assert torch.equal(torch.nan_to_num(a[key].float()), torch.nan_to_num(b[key].float())), key


def _properties(out, N):
    w, acc = out["weights"], out["acc_map"]
    assert w.shape[1] == N
    assert float(w.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-4
    assert torch.allclose(w.sum(-1), acc, atol=1e-4)
    assert bool((out["z_vals"][:, 1:] >= out["z_vals"][:, :-1]).all())
    assert torch.isfinite(out["rgb_map"]).all() and float(out["rgb_map"].min()) >= 0 and float(out["rgb_map"].max()) <= 1 + 1e-4


def test_cfg2_full_frame_properties_and_chunk_invariance():
    cfg = PN.make_cfg("cfg2")
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    batch = {k: v.to(DEV) for k, v in S.make_batch(cfg).items()}
    out = PN.make_renderer(cfg, net).render(batch)
    assert out["rgb_map"].shape == (376 * 1408, 3)
    _properties(out, 64)
    part = PN.make_renderer(PN.make_cfg("cfg2", gpu_chunk=100_003), net).render(batch)     # ragged chunks
    for k in ("rgb_map", "depth_map", "acc_map", "weights"):
        assert torch.equal(part[k], out[k]), f"{k} depends on the chunking"
    # a ray shard rendered alone equals the corresponding slice of the whole frame (what multi-GPU relies on)
    lo, hi = parallel.shard_range(batch["rays"].shape[0], 3, 8)
    shard = PN.make_renderer(cfg, net).render(parallel.shard_batch(batch, 3, 8))
    assert torch.equal(shard["rgb_map"], out["rgb_map"][lo:hi]) and torch.equal(shard["weights"], out["weights"][lo:hi])


def test_cfg3_heads_coarse_fine_strip():
    """config 3 (45 classes, 64 instances, 64 + 128 samples) on a 48-row strip: ~68k rays x 192 fine samples."""
    cfg = PN.make_cfg("cfg3")
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    batch = {k: v.to(DEV) for k, v in S.make_batch(cfg, row0=160, rows=48).items()}
    out = PN.make_renderer(cfg, net).render(batch)
    _properties(out, 192)
    assert out["semantic_map"].shape == (48 * 1408, 45) and out["instance_map"].shape == (48 * 1408, 64)
    assert torch.isfinite(out["semantic_map"]).all() and torch.isfinite(out["instance_map"]).all()
    fs = out["fixed_semantic_map"]
    assert float(fs.min()) >= 0 and bool((fs.sum(-1) <= out["acc_map"] + 1e-4).all())


def test_cfg5_equirect_shard_192_samples():
    """config 5's per-GPU share: 128 of the 1024 panorama rows (262k rays), 192 samples, both heads."""
    cfg = PN.make_cfg("cfg5")
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    scene = {k: v.to(DEV) for k, v in S.make_batch(cfg, rows=1).items() if k != "rays"}
    pose = torch.tensor([[1.0, 0, 0, 0.0], [0, 1.0, 0, 4.0], [0, 0, 1.0, 32.0]])
    out = PN.make_renderer(cfg, net).render(dict(scene, c2w=pose, row0=448, rows=128))
    assert out["rgb_map"].shape[0] == 128 * 2048
    _properties(out, 192)


def test_empty_ray_shard_renders():
    """a rank whose shard is empty (R < world) must still produce well-shaped (0-row) outputs."""
    cfg = PN.make_cfg("cfg3", N_importance=16)
    net = S.init_network_weights(PN.make_network(cfg)).to(DEV)
    batch = {k: v.to(DEV) for k, v in S.make_batch(cfg, rows=1).items()}
    batch["rays"] = batch["rays"][:0]
    out = PN.make_renderer(cfg, net).render(batch)
    assert out["rgb_map"].shape == (0, 3) and out["semantic_map"].shape == (0, 45) and out["weights"].shape == (0, 80)
    one = dict(batch, rays=S.make_rays(cfg, rows=1)[:1].to(DEV))          # and a single ray
    out = PN.make_renderer(cfg, net).render(one)
    assert out["rgb_map"].shape == (1, 3) and torch.isfinite(out["rgb_map"]).all()
