"""CPU: the oracle against hand-derived known answers and invariants ('oracle
self-tests').  The reference has no tests or golden vectors (its source is not in the mount), so these
closed forms are what pins the oracle: PARITY UNPINNED with respect to the reference itself."""
import math

import pytest
import torch

from oracle import reference_renderer as O
from panopticnerf_b200 import make_cfg, synthetic as S


def test_embed_known_values():
    x = torch.tensor([[0.0, math.pi / 2, 1.0]])
    e = O.embed(x, 2)
    assert e.shape == (1, 15)
    exp = torch.tensor([0.0, math.pi / 2, 1.0,
                        0.0, 1.0, math.sin(1.0), 1.0, math.cos(math.pi / 2), math.cos(1.0),
                        0.0, math.sin(math.pi), math.sin(2.0), 1.0, -1.0, math.cos(2.0)])
    assert torch.allclose(e[0], exp, atol=1e-6)
    assert O.embed_dim(10) == 63 and O.embed_dim(4) == 27


def test_raw2outputs_two_samples_closed_form():
    # one ray, |d| = 2, z = (1, 3): delta0 = 2*2 = 4, delta1 = 1e10*2
    raw = torch.tensor([[[0.0, 0.0, 0.0, 0.5], [2.0, -2.0, 0.0, 0.25]]])
    z = torch.tensor([[1.0, 3.0]])
    d = torch.tensor([[0.0, 0.0, 2.0]])
    out = O.raw2outputs(raw, z, d)
    a0 = 1 - math.exp(-0.5 * 4.0)
    a1 = 1.0
    w0, w1 = a0, a1 * (1 - a0 + 1e-10)
    assert out["weights"][0].tolist() == pytest.approx([w0, w1], rel=1e-6)
    sig = lambda v: 1 / (1 + math.exp(-v))
    assert out["rgb_map"][0].tolist() == pytest.approx(
        [w0 * 0.5 + w1 * sig(2.0), w0 * 0.5 + w1 * sig(-2.0), w0 * 0.5 + w1 * 0.5], rel=1e-6)
    assert float(out["depth_map"]) == pytest.approx(w0 * 1 + w1 * 3, rel=1e-6)
    assert float(out["acc_map"]) == pytest.approx(w0 + w1, rel=1e-6)
    assert float(out["disp_map"]) == pytest.approx((w0 + w1) / (w0 + 3 * w1), rel=1e-6)
    # negative density -> relu -> fully transparent -> acc 0, disp NaN (0/0), as nerf-pytorch
    out = O.raw2outputs(torch.tensor([[[0.0, 0.0, 0.0, -1.0]] * 2]), z, d)
    assert float(out["acc_map"]) == 0.0 and math.isnan(float(out["disp_map"]))
    # white background
    out = O.raw2outputs(torch.tensor([[[0.0, 0.0, 0.0, -1.0]] * 2]), z, d, white_bkgd=True)
    assert out["rgb_map"][0].tolist() == [1.0, 1.0, 1.0]


def test_semantic_compositing_and_fixed_onehot():
    raw = torch.zeros(1, 3, 4 + 2 + 3)
    raw[0, :, 3] = 1e3                          # opaque at the first sample
    raw[0, 0, 4:6] = torch.tensor([2.0, -1.0])
    raw[0, 0, 6:9] = torch.tensor([0.5, 0.25, -4.0])
    z = torch.tensor([[1.0, 2.0, 3.0]])
    d = torch.tensor([[0.0, 0.0, 1.0]])
    sb = torch.tensor([[1, -1, 0]], dtype=torch.int32)
    out = O.raw2outputs(raw, z, d, num_classes=2, num_instances=3, sample_box=sb,
                        box_sem=torch.tensor([0, 1], dtype=torch.int32), box_inst=torch.tensor([2, 5], dtype=torch.int32))
    assert out["semantic_map"][0].tolist() == pytest.approx([2.0, -1.0], abs=1e-6)
    assert out["instance_map"][0].tolist() == pytest.approx([0.5, 0.25, -4.0], abs=1e-6)
    assert out["fixed_semantic_map"][0].tolist() == pytest.approx([0.0, 1.0], abs=1e-6)   # box 1 -> class 1
    assert out["fixed_instance_map"][0].tolist() == pytest.approx([0.0, 0.0, 0.0], abs=1e-6)  # id 5 out of range
    sm = O.raw2outputs(raw, z, d, num_classes=2, num_instances=3, sem_activation="softmax")["semantic_map"][0]
    assert sm.tolist() == pytest.approx(torch.softmax(torch.tensor([2.0, -1.0]), 0).tolist(), abs=1e-6)


def test_slab_known_answers():
    c, h, rot = torch.tensor([[0.0, 0.0, 5.0]]), torch.tensor([[1.0, 2.0, 3.0]]), torch.eye(3)[None]
    o = torch.zeros(5, 3)
    d = torch.tensor([[0.0, 0.0, 1.0],      # through the centre: t in [2, 8]
                      [0.0, 0.0, -1.0],     # pointing away: tmax < 0 -> miss
                      [1.0, 0.0, 0.0],      # parallel to the z slabs, outside them: miss (inf/NaN slabs)
                      [0.2, 0.0, 1.0],      # exits through the x face
                      [0.0, 1.0, 1.0]])     # touches the edge y = 2, z = 2 only: tmax == tmin -> strict '>' says miss
    hit, bid, tin, tout = O.intersect(o, d, c, h, rot, 2)
    assert hit.tolist() == [True, False, False, True, False]
    assert tin[0, 0].item() == 2.0 and tout[0, 0].item() == 8.0
    assert bid[:, 0].tolist() == [0, -1, -1, 0, -1] and bid[:, 1].tolist() == [-1] * 5
    assert tout[3, 0].item() == pytest.approx(5.0)
    # origin inside the box: t_in clamps to 0
    hit, bid, tin, tout = O.intersect(torch.tensor([[0.0, 0.0, 5.0]]), torch.tensor([[0.0, 0.0, 1.0]]), c, h, rot, 1)
    assert hit.item() and tin.item() == 0.0 and tout.item() == 3.0
    # rotated box (45 deg about y): the ray along +z crosses the diagonal
    a = math.pi / 4
    R = torch.tensor([[[math.cos(a), 0.0, math.sin(a)], [0.0, 1.0, 0.0], [-math.sin(a), 0.0, math.cos(a)]]])
    hit, _, tin, tout = O.intersect(torch.zeros(1, 3), torch.tensor([[0.0, 0.0, 1.0]]), torch.tensor([[0.0, 0.0, 5.0]]),
                                    torch.tensor([[1.0, 1.0, 1.0]]), R, 1)
    assert hit.item() and tin.item() == pytest.approx(5 - math.sqrt(2), rel=1e-6) and tout.item() == pytest.approx(5 + math.sqrt(2), rel=1e-6)
    # M nearest, sorted by entry depth, ties by index
    cs = torch.tensor([[0.0, 0.0, 9.0], [0.0, 0.0, 3.0], [0.0, 0.0, 6.0], [0.0, 0.0, 3.0]])
    hs = torch.ones(4, 3) * 0.5
    _, bid, tin, _ = O.intersect(torch.zeros(1, 3), torch.tensor([[0.0, 0.0, 1.0]]), cs, hs, torch.eye(3)[None].repeat(4, 1, 1), 3)
    assert bid[0].tolist() == [1, 3, 2] and tin[0].tolist() == [2.5, 2.5, 5.5]


def test_stratified_and_tags():
    near, far = torch.tensor([2.0]), torch.tensor([6.0])
    t = torch.linspace(0, 1, 5)
    z = O.stratified_z(near, far, t)
    assert z[0].tolist() == [2.0, 3.0, 4.0, 5.0, 6.0]
    zj = O.stratified_z(near, far, t, 1.0, torch.zeros(1, 5))
    assert zj[0].tolist() == [2.0, 2.5, 3.5, 4.5, 5.5]                   # u = 0 -> lower edges
    zj = O.stratified_z(near, far, t, 1.0, torch.full((1, 5), 0.5))
    assert zj[0].tolist() == [2.25, 3.0, 4.0, 5.0, 5.75]
    sb = O.tag_samples(z, torch.tensor([[7, 3]], dtype=torch.int32), torch.tensor([[2.5, 3.0]]), torch.tensor([[4.0, 6.0]]))
    assert sb[0].tolist() == [-1, 7, 7, 3, 3]                             # first containing interval wins


def test_sample_pdf_delta_and_uniform():
    bins = torch.linspace(0.0, 8.0, 9)[None]                             # 9 edges, 8 weights
    w = torch.zeros(1, 8)
    w[0, 3] = 1.0
    z_f, idx = O.sample_pdf(bins, w, 5, det=True)
    assert (z_f[0, 1:-1] >= 3.0).all() and (z_f[0, 1:-1] <= 4.0).all()   # mass sits in bin [3,4]
    z_f, idx = O.sample_pdf(bins, torch.ones(1, 8), 9, det=True)
    assert torch.allclose(z_f[0], torch.linspace(0.0, 8.0, 9), atol=1e-5)  # uniform pdf -> identity
    assert idx.dtype == torch.int64 and int(idx.max()) <= 9


def test_render_invariants():
    cfg = make_cfg("cfg1", num_classes=4, num_instances=3)
    net = S.init_network_weights(O.make_network(cfg))
    batch = S.make_batch(cfg, rows=8, row0=20, num_boxes=32)
    ren = O.make_renderer(cfg, net)
    out = ren.render(batch)
    assert float(out["weights"].min()) >= 0 and float(out["acc_map"].max()) <= 1 + 1e-5
    assert (out["z_vals"][:, 1:] >= out["z_vals"][:, :-1]).all()
    # chunk invariance of batchify_rays (bit-exact on the integer outputs, 1e-6 on floats: MKL blocking)
    cfg2 = make_cfg("cfg1", num_classes=4, num_instances=3, chunk=100)
    out2 = O.make_renderer(cfg2, net).render(batch)
    for k in out:
        if out[k].dtype.is_floating_point:
            assert torch.allclose(out[k], out2[k], atol=2e-6, equal_nan=True), k
        else:
            assert torch.equal(out[k], out2[k]), k
    # ray permutation permutes outputs
    perm = torch.randperm(batch["rays"].shape[0], generator=torch.Generator().manual_seed(0))
    out3 = ren.render(dict(batch, rays=batch["rays"][perm]))
    assert torch.equal(out3["hit_mask"], out["hit_mask"][perm]) and torch.equal(out3["z_vals"], out["z_vals"][perm])
    assert torch.allclose(out3["rgb_map"], out["rgb_map"][perm], atol=2e-6)


def test_flop_counts_match_baseline_md():
    assert O.mlp_flops_per_sample(make_cfg("cfg1")) == 55040
    assert O.mlp_flops_per_sample(make_cfg("cfg2")) == 1186816
    assert O.mlp_flops_per_sample(make_cfg("cfg3")) == 1345792


def test_sample_pdf_sum_variant():
    """VERDICT r1 (weak item 1): the oracle normalises the pdf by the running sum's last element (fully specified
    order, what the kernel reproduces bit for bit); nerf-pytorch divides by torch.sum.  Both variants are kept.
    They agree in every index except where u falls within an ulp of a cdf entry: count it, and bound it."""
    g = torch.Generator().manual_seed(3)
    R, N, Ni = 4000, 64, 128
    z = torch.sort(torch.rand(R, N, generator=g) * 30 + 0.05, -1).values
    w = torch.rand(R, N, generator=g) ** 4 * (torch.rand(R, N, generator=g) > 0.5)
    bins, wi = 0.5 * (z[:, 1:] + z[:, :-1]), w[:, 1:-1]
    za, ia = O.sample_pdf(bins, wi, Ni, det=True)
    zb, ib = O.sample_pdf(bins, wi, Ni, det=True, pdf_norm="sum")
    differ = int((ia != ib).sum())
    assert differ <= 0.001 * ia.numel(), f"{differ} of {ia.numel()} searchsorted indices differ between the variants"
    same = ia == ib
    dz = (za[same] - zb[same]).abs()
    # an ulp of the cdf is amplified by 1/(cdf[above]-cdf[below]) where a bin carries almost no mass
    assert float(torch.quantile(dz, 0.999)) < 1e-3 and float(dz.max()) < 0.5
    print(f"sample_pdf: {differ} of {ia.numel()} indices differ between pdf_norm='cumsum' and 'sum'; "
          f"depths with equal index differ by at most {float(dz.max()):.2e} (99.9 %: {float(torch.quantile(dz, 0.999)):.2e})")
