"""CPU tier: the oracle of the steps beside the render path against closed-form answers -
panoptic fusion rule, multi-resolution hash-grid encoding."""
import numpy as np
import torch

from oracle import reference_panoptic as OP


def test_fusion_rule_cases():
    #            class:  0 road(stuff) 1 car(thing) 2 person(thing)
    sem = torch.tensor([[0.9, 0.1, 0.0],      # stuff
                        [0.1, 0.8, 0.1],      # car, best car slot is 1
                        [0.1, 0.2, 0.7],      # person, but no slot of that class -> falls back to stuff-like id
                        [0.4, 0.4, 0.2],      # tie -> lowest class (road)
                        [float("nan"), 0.3, 0.1]])
    inst = torch.tensor([[0.9, 0.0, 0.0], [0.2, 0.7, 0.9], [0.5, 0.4, 0.9], [0.1, 0.9, 0.0], [0.6, 0.6, 0.0]])
    is_thing, inst_class, inst_id, class_id = [0, 1, 1], [1, 1, 0], [26001, 26002, 7000], [7, 26, 24]
    pal = torch.tensor([[128, 64, 128], [0, 0, 142], [220, 20, 60]], dtype=torch.uint8)
    pan, s, k, col = OP.panoptic_fuse(sem, inst, is_thing, inst_class, inst_id, class_id, pal)
    assert s.tolist() == [0, 1, 2, 0, 1]
    assert k.tolist() == [-1, 1, -1, -1, 0]           # row 1: slot 2 has the larger score but is not a car; row 4: tie -> slot 0
    assert pan.tolist() == [7000, 26002, 24000, 7000, 26001]
    assert col[0].tolist() == [128, 64, 128] and col[2].tolist() == [220, 20, 60]
    h = (26002 * 2654435761) & 0xFFFFFFFF
    assert col[1].tolist() == [(0 + ((h >> 8) & 255) + 1) >> 1, (0 + ((h >> 16) & 255) + 1) >> 1, (142 + ((h >> 24) & 255) + 1) >> 1]
    pan2, _, k2, _ = OP.panoptic_fuse(sem, inst, is_thing, inst_class)           # no id tables: channel*1000 (+ slot + 1)
    assert pan2.tolist() == [0, 1002, 2000, 0, 1001] and k2.tolist() == k.tolist()
    pan3, _, k3, _ = OP.panoptic_fuse(sem, None, is_thing, None, None, class_id)
    assert pan3.tolist() == [7000, 26000, 24000, 7000, 26000] and (k3 == -1).all()


def test_hashgrid_oracle_properties():
    g = torch.Generator().manual_seed(0)
    L, T_log2, F = 6, 12, 2
    table = torch.randn(L, 1 << T_log2, F, generator=g)
    res = OP.hashgrid_resolutions(L, 4.0, 2.0)
    assert res == [4, 8, 16, 32, 64, 128]
    # at a grid vertex of a dense level the feature is the table entry of that vertex
    v = torch.tensor([[1, 2, 3], [4, 0, 2], [0, 0, 0]], dtype=torch.float32)
    out = OP.hashgrid_encode(v / 4.0, None, table, 4.0, 2.0)
    idx = (v[:, 0] + v[:, 1] * 5 + v[:, 2] * 25).long()
    assert torch.equal(out[:, :F], table[0][idx])
    # dense while (res+1)^3 <= T: levels 0 (125), 1 (729) are dense at T = 4096, level 2 (4913) hashes
    x = torch.tensor([[3, 5, 7]], dtype=torch.float32) / 16.0
    h = (3 ^ ((5 * 2654435761) & 0xFFFFFFFF) ^ ((7 * 805459861) & 0xFFFFFFFF)) & 4095
    assert torch.equal(OP.hashgrid_encode(x, None, table, 4.0, 2.0)[0, 2 * F:3 * F], table[2][h])
    # trilinear: along an edge of a cell the feature is linear in the coordinate
    a, b = torch.tensor([[0.25, 0.5, 0.25]]), torch.tensor([[0.5, 0.5, 0.25]])
    fa, fb = OP.hashgrid_encode(a, None, table, 4.0, 2.0)[:, :F], OP.hashgrid_encode(b, None, table, 4.0, 2.0)[:, :F]
    fm = OP.hashgrid_encode(0.7 * a + 0.3 * b, None, table, 4.0, 2.0)[:, :F]
    assert torch.allclose(fm, 0.7 * fa + 0.3 * fb, atol=1e-6)
    # aabb normalisation and clamping: outside points take the border values
    aabb = torch.tensor([[-2.0, -2.0, -2.0], [2.0, 2.0, 2.0]])
    p = (torch.rand(50, 3, generator=g) * 4 - 2)
    assert torch.equal(OP.hashgrid_encode(p, aabb, table, 4.0, 2.0), OP.hashgrid_encode((p + 2) / 4, None, table, 4.0, 2.0))
    far = torch.tensor([[9.0, 0.0, -9.0]])
    assert torch.equal(OP.hashgrid_encode(far, aabb, table, 4.0, 2.0), OP.hashgrid_encode(torch.tensor([[2.0, 0.0, -2.0]]), aabb, table, 4.0, 2.0))
    assert OP.hashgrid_encode(p, aabb, table, 4.0, 2.0).shape == (50, L * F)


def test_losses_oracle_closed_forms():
    from oracle import reference_losses as OL
    rgb, gt = torch.tensor([[0.5, 0.5, 0.5], [1.0, 0.0, 0.0]]), torch.tensor([[0.0, 0.5, 1.0], [1.0, 0.0, 0.0]])
    depth, depth_gt = torch.tensor([10.0, 20.0]), torch.tensor([12.0, 0.0])           # second ray has no stereo depth
    sem = torch.tensor([[0.0, 0.0], [2.0, 0.0]])
    fix = torch.tensor([[0.25, 0.5], [0.0, 1.0]])
    label = torch.tensor([1, 0])
    tot, t = OL.losses(rgb, None, depth, sem, fix, gt, depth_gt, label, None, (1.0, 1.0, 1.0, 1.0), False, 1e-4)
    assert float(t[0]) == np.float32(0.5 / 6) and float(t[1]) == 2.0
    assert abs(float(t[2]) - 0.5 * (np.log(2.0) + np.log1p(np.exp(-2.0)))) < 1e-6
    assert abs(float(t[3]) - 0.5 * (-np.log(0.5) - np.log(1e-4))) < 1e-5               # p = 0 is clamped to eps
    assert abs(float(tot) - float(t.sum())) < 1e-6
    _, t = OL.losses(rgb, rgb, None, sem, None, gt, None, torch.tensor([-1, 5]))        # ignored / out-of-range labels
    assert float(t[2]) == 0.0 and float(t[0]) == np.float32(1.0 / 6)
    _, t = OL.losses(None, None, None, torch.tensor([[0.2, 0.3]]), None, None, None, torch.tensor([1]), torch.tensor([0.5]), sem_is_prob=True)
    assert abs(float(t[2]) + 0.5 * np.log(0.3)) < 1e-6
