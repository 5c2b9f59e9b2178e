"""KITTI-360 3D bounding-box annotations (data_3d_bboxes/*/<sequence>.xml) -> the primitive table of the render
path (center [B,3], half extents [B,3], rotation [B,3,3] with the box axes as columns, class id, instance id).

XML layout (public annotation format, recalled): <opencv_storage> holds one <objectN> per box with
  <transform type_id="opencv-matrix"> rows 4, cols 4, data = 16 floats, row-major  (box frame -> world: [A | T])
  <vertices  type_id="opencv-matrix"> rows 8 (or more for meshes), cols 3, in the box frame
  <semanticId>, <instanceId>, <timestamp> (-1: static, else the frame a dynamic box belongs to), <dynamic>
A = R diag(s): the annotation's cuboids are axis-aligned boxes of the box frame, rotated and scaled into the world."""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional

import numpy as np


@dataclass
class Box3D:
    name: str
    transform: np.ndarray      # 4x4, box frame -> world
    vertices: np.ndarray       # [V,3] in the box frame
    semantic_id: int
    instance_id: int
    timestamp: int = -1        # -1 = static
    dynamic: int = 0

    def world_vertices(self) -> np.ndarray:
        return self.vertices @ self.transform[:3, :3].T + self.transform[:3, 3]


def _matrix(node, what: str) -> np.ndarray:
    rows, cols = int(node.findtext("rows")), int(node.findtext("cols"))
    data = np.array(node.findtext("data").split(), dtype=np.float64)
    if data.size != rows * cols:
        raise ValueError(f"{what}: {data.size} values for a {rows}x{cols} matrix")
    return data.reshape(rows, cols)


def parse_bboxes_xml(path) -> List[Box3D]:
    root = ET.parse(str(path)).getroot()
    out = []
    for obj in root:
        if obj.find("transform") is None:
            continue
        tr, vt = _matrix(obj.find("transform"), f"{obj.tag}/transform"), _matrix(obj.find("vertices"), f"{obj.tag}/vertices")
        if tr.shape != (4, 4) or vt.shape[1] != 3:
            raise ValueError(f"{path}: {obj.tag}: transform {tr.shape}, vertices {vt.shape}")
        gi = lambda k, d: int(float(obj.findtext(k))) if obj.findtext(k) not in (None, "") else d
        out.append(Box3D(obj.tag, tr, vt, gi("semanticId", -1), gi("instanceId", -1), gi("timestamp", -1), gi("dynamic", 0)))
    return out


def boxes_to_primitives(boxes: Iterable[Box3D], frame: Optional[int] = None, ortho_tol: float = 1e-3) -> Dict[str, np.ndarray]:
    """Static boxes (+ the dynamic ones stamped `frame`) as oriented cuboids.  Per box, with [lo, hi] the bounds of
    its vertices in the box frame and A the 3x3 of its transform: s_j = |A[:, j]|, rot = A / s (a reflection is
    removed by flipping the last axis - a cuboid does not care), half = (hi - lo)/2 * s, center = A (lo + hi)/2 + T.
    A box whose axes are not orthogonal within `ortho_tol` is rejected (the slab test needs a rotation)."""
    c, h, r, sem, inst, names = [], [], [], [], [], []
    for b in boxes:
        if b.timestamp != -1 and (frame is None or b.timestamp != frame):
            continue
        A, T = b.transform[:3, :3], b.transform[:3, 3]
        s = np.linalg.norm(A, axis=0)
        if np.any(s <= 0):
            raise ValueError(f"{b.name}: degenerate transform")
        R = A / s
        if np.abs(R.T @ R - np.eye(3)).max() > ortho_tol:
            raise ValueError(f"{b.name}: transform axes are not orthogonal (max |R^T R - I| = {np.abs(R.T @ R - np.eye(3)).max():.2e})")
        if np.linalg.det(R) < 0:
            R = R * np.array([1.0, 1.0, -1.0])
        lo, hi = b.vertices.min(0), b.vertices.max(0)
        c.append(A @ ((lo + hi) * 0.5) + T)
        h.append((hi - lo) * 0.5 * s)
        r.append(R)
        sem.append(b.semantic_id)
        inst.append(b.instance_id)
        names.append(b.name)
    n = len(c)
    return {"box_center": np.asarray(c, dtype=np.float32).reshape(n, 3), "box_half": np.asarray(h, dtype=np.float32).reshape(n, 3),
            "box_rot": np.asarray(r, dtype=np.float32).reshape(n, 3, 3), "box_sem": np.asarray(sem, dtype=np.int32),
            "box_inst": np.asarray(inst, dtype=np.int32), "names": names}


def primitive_batch(prims: Dict[str, np.ndarray], sem_to_train: Optional[Dict[int, int]] = None,
                    inst_to_slot: Optional[Dict[int, int]] = None, device="cuda") -> Dict[str, "object"]:
    """The primitive block of a render batch (torch tensors on `device`).  `sem_to_train` maps KITTI-360 semanticIds
    to the network's class channels, `inst_to_slot` instanceIds to its instance channels; boxes whose id has no
    channel get -1 (they still bound samples, they just do not vote in the fixed maps)."""
    import torch
    sem = np.array([(sem_to_train or {}).get(int(s), int(s) if sem_to_train is None else -1) for s in prims["box_sem"]], dtype=np.int32)
    inst = np.array([(inst_to_slot or {}).get(int(s), int(s) if inst_to_slot is None else -1) for s in prims["box_inst"]], dtype=np.int32)
    t = lambda a, dt: torch.as_tensor(a, dtype=dt).to(device)
    return {"box_center": t(prims["box_center"], torch.float32), "box_half": t(prims["box_half"], torch.float32),
            "box_rot": t(prims["box_rot"], torch.float32), "box_sem": t(sem, torch.int32), "box_inst": t(inst, torch.int32)}
