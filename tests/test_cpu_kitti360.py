"""CPU tier: the KITTI-360 on-disk formats in front of the render path - calibration / pose
readers, the fisheye yaml and ray model, bounding-box XML -> primitive table, intersection cache.  Fixtures are
written by the tests in the documented layouts (the dataset itself is not available here)."""
import numpy as np
import pytest
import torch

from oracle import reference_renderer as O
from panopticnerf_b200.lib.datasets import kitti360 as K

FISHEYE = dict(xi=2.2134047507854890, k1=1.6798235660113681e-02, k2=1.6548773243373522, p1=4.2e-04, p2=4.2e-04,
               gamma1=1.3363220825849971e+03, gamma2=1.3357883350012958e+03, u0=7.1694323510126321e+02, v0=7.0576498308221585e+02)


def _rot(ax, ang):
    ax = np.asarray(ax, dtype=np.float64) / np.linalg.norm(ax)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def test_calibration_and_pose_readers(tmp_path):
    (tmp_path / "perspective.txt").write_text(
        "calib_time: 09-Jan-2020\nS_rect_00: 1408 376\n"
        "P_rect_00: 552.554261 0.000000 682.049453 0.000000 0.000000 552.554261 238.769549 0.000000 0.000000 0.000000 1.000000 0.000000\n"
        "R_rect_00: 0.999974 -0.007141 -0.000089 0.007141 0.999969 -0.003247 0.000112 0.003247 0.999995\n"
        "P_rect_01: 552.554261 0.000000 682.049453 -328.318735 0.000000 552.554261 238.769549 0.000000 0.000000 0.000000 1.000000 0.000000\n")
    p = K.load_perspective(tmp_path / "perspective.txt")
    assert p["size"] == (1408, 376) and p["K"][0, 0] == pytest.approx(552.554261) and p["K"][1, 2] == pytest.approx(238.769549)
    assert K.load_perspective(tmp_path / "perspective.txt", cam=1)["P_rect"][0, 3] == pytest.approx(-328.318735)
    with pytest.raises(ValueError, match="P_rect_02"):
        K.load_perspective(tmp_path / "perspective.txt", cam=2)
    (tmp_path / "calib_cam_to_pose.txt").write_text(
        "image_00: 0.0371783278 -0.0986182135 0.9944306009 1.5752681039 0.9992675562 -0.0053553387 -0.0378902567 0.0043914093 "
        "0.0090621821 0.9951109327 0.0983468786 -0.6500000000\nimage_02: 1 0 0 0.7 0 1 0 0.1 0 0 1 -0.6\n")
    c2p = K.load_cam_to_pose(tmp_path / "calib_cam_to_pose.txt")
    assert c2p["image_00"].shape == (4, 4) and c2p["image_00"][0, 3] == pytest.approx(1.5752681039) and c2p["image_02"][3, 3] == 1.0
    pose = np.eye(4); pose[:3, :3] = _rot([0, 0, 1], 0.3); pose[:3, 3] = [100.0, -20.0, 3.0]
    (tmp_path / "cam0_to_world.txt").write_text("7 " + " ".join(f"{x:.9f}" for x in pose.reshape(-1)) + "\n\n9 " +
                                                " ".join(f"{x:.9f}" for x in np.eye(4).reshape(-1)) + "\n")
    c2w = K.load_cam0_to_world(tmp_path / "cam0_to_world.txt")
    assert sorted(c2w) == [7, 9] and np.allclose(c2w[7], pose, atol=1e-8)
    (tmp_path / "poses.txt").write_text("7 " + " ".join(f"{x:.9f}" for x in pose[:3].reshape(-1)) + "\n")
    assert np.allclose(K.load_poses(tmp_path / "poses.txt")[7], pose, atol=1e-8)
    (tmp_path / "bad.txt").write_text("7 1 2 3\n")
    with pytest.raises(ValueError, match="expected 16"):
        K.load_cam0_to_world(tmp_path / "bad.txt")
    b = K.perspective_batch(p, c2w[7])
    assert b["intrinsics"] == pytest.approx((552.554261, 552.554261, 682.049453, 238.769549)) and b["c2w"].shape == (3, 4)


def _write_fisheye_yaml(path, fe=FISHEYE, model="MEI"):
    path.write_text(f"%YAML:1.0\n---\nmodel_type: {model}\ncamera_name: image_02\nimage_width: 1400\nimage_height: 1400\n"
                    f"mirror_parameters:\n   xi: {fe['xi']:.16e}\ndistortion_parameters:\n   k1: {fe['k1']:.16e}\n   k2: {fe['k2']:.16e}\n"
                    f"   p1: {fe['p1']:.16e}\n   p2: {fe['p2']:.16e}\nprojection_parameters:\n   gamma1: {fe['gamma1']:.16e}\n"
                    f"   gamma2: {fe['gamma2']:.16e}\n   u0: {fe['u0']:.16e}\n   v0: {fe['v0']:.16e}\n")


def test_fisheye_yaml_and_ray_model_round_trip(tmp_path):
    _write_fisheye_yaml(tmp_path / "image_02.yaml")
    fe = K.load_fisheye_yaml(tmp_path / "image_02.yaml")
    assert fe["xi"] == pytest.approx(FISHEYE["xi"]) and fe["image_width"] == 1400 and fe["v0"] == pytest.approx(FISHEYE["v0"])
    _write_fisheye_yaml(tmp_path / "pinhole.yaml", model="PINHOLE")
    with pytest.raises(ValueError, match="MEI"):
        K.load_fisheye_yaml(tmp_path / "pinhole.yaml")
    # rays of every 25th pixel of rows 100..1300, identity pose; project them back with the published forward model
    # (kitti360scripts CameraFisheye.cam2image without the tangential terms) in float64
    blk = K.fisheye_batch(fe, np.eye(4))
    H = W = 1400
    rays = O.generate_rays(H, W, blk["intrinsics"], blk["c2w"], "fisheye", row0=100, rows=1200).reshape(1200, W, 6)[::25, ::25]
    d = rays[..., 3:].double().reshape(-1, 3)
    assert torch.all(rays[..., :3] == 0) and torch.isfinite(d).all()      # (also beyond the field of view: clamped)
    v, u = torch.meshgrid(torch.arange(100, 1300, 25, dtype=torch.float64), torch.arange(0, W, 25, dtype=torch.float64), indexing="ij")
    n = d / d.norm(dim=-1, keepdim=True)
    x, y = n[:, 0] / (n[:, 2] + fe["xi"]), n[:, 1] / (n[:, 2] + fe["xi"])
    r2 = x * x + y * y
    rad = 1 + fe["k1"] * r2 + fe["k2"] * r2 * r2
    uu, vv = fe["gamma1"] * x * rad + fe["u0"], fe["gamma2"] * y * rad + fe["v0"]
    inside = ((u.reshape(-1) - fe["u0"]) ** 2 + (v.reshape(-1) - fe["v0"]) ** 2).sqrt() < 690       # the image circle
    assert inside.sum() > 1500
    assert torch.allclose(d.norm(dim=-1)[inside], torch.ones(int(inside.sum()), dtype=torch.float64), atol=1e-5)   # unit sphere
    assert (uu - u.reshape(-1))[inside].abs().max() < 2e-3 and (vv - v.reshape(-1))[inside].abs().max() < 2e-3   # pixels
    # the principal point looks along +z; a pose rotates the directions and sets the origin
    c = O.generate_rays(3, 3, (100.0, 100.0, 1.0, 1.0, fe["xi"], fe["k1"], fe["k2"]), torch.eye(4)[:3], "fisheye")[4]
    assert torch.allclose(c[3:], torch.tensor([0.0, 0.0, 1.0]), atol=1e-7)
    pose = np.eye(4); pose[:3, :3] = _rot([0, 1, 0], np.pi / 2); pose[:3, 3] = [1, 2, 3]
    c = O.generate_rays(3, 3, (100.0, 100.0, 1.0, 1.0, fe["xi"], fe["k1"], fe["k2"]), torch.tensor(pose[:3], dtype=torch.float32), "fisheye")[4]
    assert torch.allclose(c, torch.tensor([1.0, 2.0, 3.0, 1.0, 0.0, 0.0]), atol=1e-6)


def _mat(name, a):
    a = np.asarray(a, dtype=np.float64)
    return (f'<{name} type_id="opencv-matrix"><rows>{a.shape[0]}</rows><cols>{a.shape[1]}</cols><dt>f</dt><data>\n' +
            " ".join(f"{x:.8e}" for x in a.reshape(-1)) + f"</data></{name}>")


def _obj(i, transform, vertices, sem, inst, timestamp=-1):
    return (f"<object{i}><index>{i}</index><label>x</label>{_mat('transform', transform)}{_mat('vertices', vertices)}"
            f"<semanticId>{sem}</semanticId><instanceId>{inst}</instanceId><timestamp>{timestamp}</timestamp>"
            f"<dynamic>{int(timestamp != -1)}</dynamic></object{i}>")


CUBE = np.array([[x, y, z] for x in (-0.5, 0.5) for y in (-0.5, 0.5) for z in (-0.5, 0.5)])


def _tf(R, s, T):
    m = np.eye(4); m[:3, :3] = R * np.asarray(s); m[:3, 3] = T
    return m


def test_bbox_xml_to_primitive_table(tmp_path):
    R1, R2 = _rot([0.2, 0.1, 1.0], 0.7), _rot([1.0, 0.0, 0.3], -1.1)
    refl = R2 * np.array([1.0, -1.0, 1.0])                               # a left-handed annotation frame
    objs = [_obj(1, _tf(R1, [4.0, 2.0, 1.5], [10, 20, 1]), CUBE, 26, 26001),
            _obj(2, _tf(refl, [1.0, 3.0, 2.0], [-5, 4, 0.5]), CUBE * [2.0, 1.0, 1.0] + [1.0, 0.0, 0.5], 11, 11007),
            _obj(3, _tf(np.eye(3), [2, 2, 2], [0, 0, 0]), CUBE, 26, 26002, timestamp=40)]
    (tmp_path / "seq.xml").write_text("<?xml version=\"1.0\"?>\n<opencv_storage>" + "".join(objs) + "</opencv_storage>")
    boxes = K.parse_bboxes_xml(tmp_path / "seq.xml")
    assert [b.semantic_id for b in boxes] == [26, 11, 26] and boxes[2].timestamp == 40 and boxes[1].vertices.shape == (8, 3)
    static = K.boxes_to_primitives(boxes)
    assert static["box_center"].shape == (2, 3) and static["names"] == ["object1", "object2"]
    frame40 = K.boxes_to_primitives(boxes, frame=40)
    assert frame40["box_center"].shape == (3, 3) and list(frame40["box_inst"]) == [26001, 11007, 26002]
    assert np.allclose(static["box_half"][0], [2.0, 1.0, 0.75]) and np.allclose(static["box_center"][0], [10, 20, 1])
    assert np.allclose(static["box_half"][1], [1.0, 1.5, 1.0], atol=1e-6)
    rng = np.random.default_rng(0)
    for k, b in enumerate(boxes[:2]):
        c, h, r = static["box_center"][k].astype(np.float64), static["box_half"][k].astype(np.float64), static["box_rot"][k].astype(np.float64)
        assert np.allclose(r.T @ r, np.eye(3), atol=1e-6) and np.linalg.det(r) > 0
        loc = (b.world_vertices() - c) @ r                              # the annotation's corners are the cuboid's corners
        assert np.allclose(np.abs(loc), np.broadcast_to(h, (8, 3)), atol=1e-5)
        lo, hi = b.vertices.min(0), b.vertices.max(0)
        p_in = (rng.uniform(lo, hi, size=(200, 3))) @ b.transform[:3, :3].T + b.transform[:3, 3]
        assert np.all(np.abs((p_in - c) @ r) <= h + 1e-5)
    # through the oracle's slab test: a ray aimed at box 1's centre enters and leaves half a box-diagonal-chord apart
    t = lambda a: torch.tensor(a, dtype=torch.float32)
    o = t([[0.0, 0.0, 1.0]]); d = t(static["box_center"][:1]) - o; d = d / d.norm()
    hit, box_id, t_in, t_out = O.intersect(o, d, t(static["box_center"]), t(static["box_half"]), t(static["box_rot"]), 2)
    assert bool(hit[0]) and int(box_id[0, 0]) == 0
    mid = 0.5 * (t_in[0, 0] + t_out[0, 0])
    assert float(mid) == pytest.approx(float((t(static["box_center"][0]) - o[0]).norm()), rel=1e-5)
    # a sheared annotation is refused
    shear = np.eye(4); shear[0, 1] = 0.4
    (tmp_path / "bad.xml").write_text("<opencv_storage>" + _obj(1, shear, CUBE, 1, 1) + "</opencv_storage>")
    with pytest.raises(ValueError, match="orthogonal"):
        K.boxes_to_primitives(K.parse_bboxes_xml(tmp_path / "bad.xml"))
    blk = K.primitive_batch(static, sem_to_train={26: 2}, inst_to_slot={26001: 0, 11007: 5}, device="cpu")
    assert blk["box_sem"].tolist() == [2, -1] and blk["box_inst"].tolist() == [0, 5] and blk["box_rot"].shape == (2, 3, 3)
    assert K.boxes_to_primitives([])["box_center"].shape == (0, 3)


def test_intersection_cache_round_trip_and_staleness(tmp_path):
    rng = np.random.default_rng(1)
    prims = {"box_center": rng.normal(size=(5, 3)).astype(np.float32), "box_half": rng.uniform(0.5, 2, (5, 3)).astype(np.float32),
             "box_rot": np.stack([_rot(rng.normal(size=3), 0.4) for _ in range(5)]).astype(np.float32)}
    H, W, M = 6, 10, 4
    c2w, intr = np.eye(4)[:3], (50.0, 50.0, 5.0, 3.0)
    rays = O.generate_rays(H, W, intr, torch.tensor(c2w, dtype=torch.float32))
    t = lambda a: torch.tensor(a)
    hit, box_id, t_in, t_out = O.intersect(rays[:, :3] + torch.tensor([0.0, 0.0, -8.0]), rays[:, 3:], t(prims["box_center"]),
                                           t(prims["box_half"]), t(prims["box_rot"]), M)
    K.save_intersections(tmp_path / "f.npz", hit, box_id, t_in, t_out, prims, c2w, intr, H, W)
    got = K.load_intersections(tmp_path / "f.npz", prims, c2w, intr, H, W)
    assert np.array_equal(got["hit_mask"], hit.numpy()) and np.array_equal(got["box_id"], box_id.numpy())
    assert np.array_equal(got["t_in"], t_in.numpy()) and np.array_equal(got["t_out"], t_out.numpy())
    moved = dict(prims, box_center=prims["box_center"] + 1e-3)
    with pytest.raises(ValueError, match="stale"):
        K.load_intersections(tmp_path / "f.npz", moved, c2w, intr, H, W)
