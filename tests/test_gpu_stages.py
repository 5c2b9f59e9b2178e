"""GPU parity, stage by stage, CUDA (through the libpnr C ABI) vs the oracle on identical inputs.
Integer / mask / index outputs: bit-exact.  Floating point: 1e-4 relative with the floors of SURVEY 8(a).
Parity is "vs in-repo oracle" - the reference source is not in the mount (parity unpinned)."""
import math

import pytest
import torch

from oracle import reference_renderer as O
from panopticnerf_b200 import make_cfg, synthetic as S
from panopticnerf_b200.lib.networks.renderer import panopticnerf_renderer as P
from util import assert_close, rms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rays(cfg, rows=8):
    return S.make_rays(cfg, seed=0, row0=100, rows=rows)


@pytest.mark.parametrize("B,M", [(64, 4), (7, 1), (300, 8), (1, 4), (600, 3)])
def test_intersect_bit_exact(B, M):
    cfg = make_cfg("cfg2")
    rays = _rays(cfg)
    boxes = S.make_boxes(B, 45, 64, seed=3)
    hit, bid, tin, tout = O.intersect(rays[:, :3], rays[:, 3:], boxes["box_center"], boxes["box_half"],
                                      boxes["box_rot"], M)
    g = P.intersect(rays.to(DEV), boxes["box_center"].to(DEV), boxes["box_half"].to(DEV),
                    boxes["box_rot"].to(DEV), M)
    assert torch.equal(g[0].cpu(), hit)
    assert torch.equal(g[1].cpu(), bid)
    assert torch.equal(g[2].cpu(), tin), (g[2].cpu() - tin).abs().max()
    assert torch.equal(g[3].cpu(), tout)
    assert B < 64 or hit.any()
    assert B != 64 or not hit.all()


def test_intersect_edge_cases():
    """axis-parallel rays (division by zero -> inf/NaN slabs), origin inside a box, corner grazing."""
    c = torch.tensor([[0.0, 0.0, 5.0], [0.0, 0.0, 0.0], [3.0, 0.0, 5.0]])
    h = torch.tensor([[1.0, 1.0, 1.0], [0.5, 0.5, 0.5], [1.0, 2.0, 1.0]])
    rot = torch.eye(3)[None].repeat(3, 1, 1)
    o = torch.tensor([[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0], [2.0, 0.0, 0.0],
                      [0.0, 0.0, 10.0]])
    d = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0],
                      [0.0, 0.0, 1.0]])
    ref = O.intersect(o, d, c, h, rot, 3)
    g = P.intersect(torch.cat([o, d], -1).to(DEV), c.to(DEV), h.to(DEV), rot.to(DEV), 3)
    for a, b in zip(g, ref):
        assert torch.equal(a.cpu(), b)


def test_intersect_empty_inputs():
    cfg = make_cfg("cfg2")
    boxes = S.make_boxes(4, 45, 64)
    g = P.intersect(torch.zeros(0, 6, device=DEV), boxes["box_center"].to(DEV), boxes["box_half"].to(DEV),
                    boxes["box_rot"].to(DEV), 4)
    assert g[0].shape == (0,) and g[1].shape == (0, 4)


def test_scene_near_far_and_bound():
    cfg = make_cfg("cfg2")
    rays = _rays(cfg)
    aabb = torch.tensor(S.SCENE_AABB)
    near, far = O.scene_near_far(rays[:, :3], rays[:, 3:], aabb, cfg.near, cfg.far)
    gn, gf = P.scene_near_far(rays.to(DEV), aabb, cfg.near, cfg.far)
    assert torch.equal(gn.cpu(), near) and torch.equal(gf.cpu(), far)
    # rays that miss the scene box fall back to (near_min, far_default)
    o = torch.tensor([[100.0, 0.0, 0.0]]).repeat(4, 1)
    d = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.1], [0.0, 0.0, 1.0]])
    near2, far2 = O.scene_near_far(o, d, aabb, cfg.near, cfg.far)
    gn2, gf2 = P.scene_near_far(torch.cat([o, d], -1).to(DEV), aabb, cfg.near, cfg.far)
    assert torch.equal(gn2.cpu(), near2) and torch.equal(gf2.cpu(), far2)
    # bound_by_primitives
    boxes = S.make_boxes(64, 45, 64)
    hit, bid, tin, tout = O.intersect(rays[:, :3], rays[:, 3:], boxes["box_center"], boxes["box_half"],
                                      boxes["box_rot"], 4)
    first = tin[:, 0]
    last = torch.where(bid >= 0, tout, torch.zeros_like(tout)).max(1).values
    n_ref = torch.where(hit, torch.maximum(near, first), near)
    f_ref = torch.where(hit, torch.minimum(far, last), far)
    n_g, f_g = P.bound_by_primitives(hit.to(DEV), bid.to(DEV), tin.to(DEV), tout.to(DEV), near.to(DEV), far.to(DEV))
    assert torch.equal(n_g.cpu(), n_ref) and torch.equal(f_g.cpu(), f_ref)


@pytest.mark.parametrize("N,perturb", [(64, 0.0), (64, 1.0), (32, 1.0), (192, 0.0), (1, 0.0), (7, 1.0)])
def test_stratified_and_tags_bit_exact(N, perturb):
    cfg = make_cfg("cfg2")
    rays = _rays(cfg, rows=4)
    aabb = torch.tensor(S.SCENE_AABB)
    near, far = O.scene_near_far(rays[:, :3], rays[:, 3:], aabb, cfg.near, cfg.far)
    boxes = S.make_boxes(64, 45, 64)
    hit, bid, tin, tout = O.intersect(rays[:, :3], rays[:, 3:], boxes["box_center"], boxes["box_half"],
                                      boxes["box_rot"], 4)
    t_vals = torch.linspace(0.0, 1.0, N)
    u = torch.rand(rays.shape[0], N, generator=torch.Generator().manual_seed(5))
    z = O.stratified_z(near, far, t_vals, perturb, u)
    sb = O.tag_samples(z, bid, tin, tout)
    gz, gsb = P.stratified_z(near.to(DEV), far.to(DEV), t_vals.to(DEV), perturb, u.to(DEV), bid.to(DEV),
                             tin.to(DEV), tout.to(DEV), want_tags=True)
    assert torch.equal(gz.cpu(), z), float((gz.cpu() - z).abs().max())
    assert torch.equal(gsb.cpu(), sb)
    assert torch.equal(P.tag_samples(gz, bid.to(DEV), tin.to(DEV), tout.to(DEV)).cpu(), sb)
    if N >= 32:
        assert (sb >= 0).any()


@pytest.mark.parametrize("N,Ni,det", [(64, 128, True), (64, 128, False), (32, 16, True), (3, 5, True), (192, 64, False)])
def test_sample_pdf_indices_bit_exact(N, Ni, det):
    g = torch.Generator().manual_seed(11)
    R = 3000
    z = torch.sort(torch.rand(R, N, generator=g) * 30 + 0.05, -1).values
    w = torch.rand(R, N, generator=g) ** 4
    w[: R // 4] *= (torch.rand(R // 4, N, generator=g) > 0.8)          # sparse weights: flat cdf runs
    w[R // 4: R // 4 + 50] = 0.0                                        # all-zero rows: uniform pdf
    w = w / (w.sum(-1, keepdim=True) + 1e-3)
    u = None if det else torch.rand(R, Ni, generator=g)
    zm = 0.5 * (z[:, 1:] + z[:, :-1])
    z_f, idx = O.sample_pdf(zm, w[:, 1:-1], Ni, det=det, u=u)
    z_all = O.merge_sorted(z, z_f)
    gz_f, gz_all, gidx = P.sample_pdf(z.to(DEV), w.to(DEV), Ni, det=det, u=None if u is None else u.to(DEV),
                                      want_idx=True)
    assert torch.equal(gidx.cpu(), idx), int((gidx.cpu() != idx).sum())
    assert torch.equal(gz_f.cpu(), z_f), float((gz_f.cpu() - z_f).abs().max())
    assert torch.equal(gz_all.cpu(), z_all)
    assert bool((gz_all[:, 1:] >= gz_all[:, :-1]).all())                # sortedness property


@pytest.mark.parametrize("L", [10, 4, 0])
def test_encode(L):
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(1000, 3, generator=g) - 0.5) * 128.0
    x[0] = 0.0
    x[1] = math.pi / 2
    ref = O.embed(x, L)
    out = P.embed(x.to(DEV), L)
    assert out.shape == ref.shape
    assert_close(out, ref, 1.0, f"embed L={L}", rel=2e-6)             # sin/cos values are O(1); ulp-level
    assert torch.equal(out[:, :3].cpu(), x)


@pytest.mark.parametrize("N,C,K,flags", [(64, 0, 0, {}), (64, 45, 64, {}), (32, 5, 6, {"white_bkgd": True}),
                                         (192, 45, 64, {"sem_activation": "softmax"}),
                                         (48, 45, 0, {"mask_outside": True}), (1, 0, 0, {}), (33, 3, 70, {})])
def test_raw2outputs(N, C, K, flags):
    g = torch.Generator().manual_seed(7)
    R = 500
    raw = torch.randn(R, N, 4 + C + K, generator=g)
    raw[..., 3] = raw[..., 3] * 0.3 + 0.05
    raw[: R // 10, :, 3] = -1.0                                        # fully transparent rays: acc = 0, disp NaN
    z = torch.sort(torch.rand(R, N, generator=g) * 40 + 0.05, -1).values
    d = torch.randn(R, 3, generator=g)
    B = 20
    sb = torch.randint(-1, B, (R, N), generator=g, dtype=torch.int32)
    bs = torch.randint(0, max(C, 1), (B,), generator=g, dtype=torch.int32)
    bi = torch.randint(0, max(K, 1), (B,), generator=g, dtype=torch.int32)
    ref = O.raw2outputs(raw, z, d, num_classes=C, num_instances=K, sample_box=sb, box_sem=bs, box_inst=bi, **flags)
    out = P.raw2outputs(raw.to(DEV), z.to(DEV), d.to(DEV), num_classes=C, num_instances=K,
                        sample_box=sb.to(DEV), box_sem=bs.to(DEV), box_inst=bi.to(DEV), **flags)
    assert set(out) == set(ref)
    floors = {"rgb_map": 1e-2, "acc_map": 1e-3, "weights": 1e-3, "depth_map": 0.4, "disp_map": 1.0 / 40}
    for k, v in ref.items():
        f = floors.get(k, max(rms(v), 1e-6))
        assert_close(out[k], v, f, k)
    assert float(out["acc_map"].max()) <= 1.0 + 1e-5 and float(out["weights"].min()) >= 0.0
