"""KITTI-360 calibration / pose readers (plain text and the OpenCV-style fisheye yaml) and the camera blocks of a
render batch.  Formats (public dataset documentation, recalled - the reference's loader is not in the mount):

  calibration/perspective.txt        "P_rect_00: <12 floats>", "R_rect_00: <9 floats>", "S_rect_00: <w> <h>", ...
  calibration/calib_cam_to_pose.txt  "image_00: <12 floats>"  (3x4 camera -> GPS/IMU pose frame), image_01..03
  calibration/image_02.yaml / _03    MEI model: mirror_parameters.xi, distortion_parameters.k1 k2 p1 p2,
                                     projection_parameters.gamma1 gamma2 u0 v0, image_width / image_height
  data_poses/<seq>/cam0_to_world.txt "<frame> <16 floats>"    (4x4 rectified camera 0 -> world)
  data_poses/<seq>/poses.txt         "<frame> <12 floats>"    (3x4 GPS/IMU pose -> world)
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Tuple

import numpy as np


def _floats(tokens, n: int, what: str) -> np.ndarray:
    if len(tokens) != n:
        raise ValueError(f"{what}: expected {n} numbers, got {len(tokens)}")
    return np.array([float(t) for t in tokens], dtype=np.float64)


def _keyed_lines(path) -> Dict[str, list]:
    out = {}
    for ln in Path(path).read_text().splitlines():
        ln = ln.strip()
        if not ln or ln.startswith("#"):
            continue
        key, sep, rest = ln.partition(":")
        if not sep:
            raise ValueError(f"{path}: line without 'key:' prefix: {ln[:60]!r}")
        out[key.strip()] = rest.split()
    return out


def load_perspective(path, cam: int = 0) -> Dict[str, np.ndarray]:
    """perspective.txt -> {'K': 3x3 of P_rect_0<cam>, 'P_rect': 3x4, 'R_rect': 3x3, 'size': (W, H)}."""
    kv = _keyed_lines(path)
    tag = f"{cam:02d}"
    if f"P_rect_{tag}" not in kv:
        raise ValueError(f"{path}: no P_rect_{tag}")
    P = _floats(kv[f"P_rect_{tag}"], 12, f"P_rect_{tag}").reshape(3, 4)
    R = _floats(kv[f"R_rect_{tag}"], 9, f"R_rect_{tag}").reshape(3, 3) if f"R_rect_{tag}" in kv else np.eye(3)
    size = tuple(int(float(t)) for t in kv[f"S_rect_{tag}"]) if f"S_rect_{tag}" in kv else None
    return {"K": P[:, :3].copy(), "P_rect": P, "R_rect": R, "size": size}


def load_cam_to_pose(path) -> Dict[str, np.ndarray]:
    """calib_cam_to_pose.txt -> {'image_00': 4x4, ...} (camera -> GPS/IMU pose frame)."""
    out = {}
    for key, toks in _keyed_lines(path).items():
        m = np.eye(4)
        m[:3, :4] = _floats(toks, 12, key).reshape(3, 4)
        out[key] = m
    return out


def _framed(path, n: int) -> Dict[int, np.ndarray]:
    out = {}
    for ln in Path(path).read_text().splitlines():
        toks = ln.split()
        if not toks:
            continue
        out[int(float(toks[0]))] = _floats(toks[1:], n, f"{path} frame {toks[0]}")
    return out


def load_cam0_to_world(path) -> Dict[int, np.ndarray]:
    """cam0_to_world.txt -> {frame: 4x4}."""
    return {f: v.reshape(4, 4) for f, v in _framed(path, 16).items()}


def load_poses(path) -> Dict[int, np.ndarray]:
    """poses.txt -> {frame: 4x4} (GPS/IMU pose -> world; the 3x4 of the file completed with [0 0 0 1])."""
    out = {}
    for f, v in _framed(path, 12).items():
        m = np.eye(4)
        m[:3, :4] = v.reshape(3, 4)
        out[f] = m
    return out


def load_fisheye_yaml(path) -> Dict[str, float]:
    """image_02.yaml / image_03.yaml (OpenCV '%YAML:1.0' file, two-level 'section:\\n   key: value').  Returns xi, k1,
    k2, p1, p2, gamma1, gamma2, u0, v0, image_width, image_height.  Only the MEI model is known here."""
    out: Dict[str, float] = {}
    model = None
    for ln in Path(path).read_text().splitlines():
        if ln.startswith("%") or not ln.strip() or ln.strip() == "---":
            continue
        key, sep, val = ln.strip().partition(":")
        if not sep:
            raise ValueError(f"{path}: cannot parse line {ln!r}")
        val = val.strip()
        if key == "model_type":
            model = val
        elif val == "" or key == "camera_name":
            continue                      # a section header (mirror_parameters: ...) or a name
        else:
            out[key] = float(val)
    if model is not None and model != "MEI":
        raise ValueError(f"{path}: model_type {model!r}; only the MEI (unified) fisheye model is implemented")
    missing = [k for k in ("xi", "k1", "k2", "gamma1", "gamma2", "u0", "v0") if k not in out]
    if missing:
        raise ValueError(f"{path}: missing {missing}")
    return out


def perspective_batch(persp: Dict[str, np.ndarray], cam_to_world: np.ndarray) -> Dict[str, object]:
    """Camera block of a render batch for the rectified perspective camera: Renderer.render generates the rays on
    the device from it (camera 'pinhole': fx, fy, cx, cy of P_rect; pose = the 3x4 of cam_to_world)."""
    import torch
    K = persp["K"]
    return {"intrinsics": (float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])),
            "c2w": torch.tensor(np.asarray(cam_to_world)[:3, :4], dtype=torch.float32)}


def fisheye_batch(fe: Dict[str, float], cam_to_world: np.ndarray) -> Dict[str, object]:
    """Camera block for a fisheye camera (camera 'fisheye' of generate_rays).  The tangential terms p1, p2 of the
    yaml (4e-4 on KITTI-360) are not part of the ray model."""
    import torch
    return {"intrinsics": (fe["gamma1"], fe["gamma2"], fe["u0"], fe["v0"], fe["xi"], fe["k1"], fe["k2"]),
            "c2w": torch.tensor(np.asarray(cam_to_world)[:3, :4], dtype=torch.float32)}


def fisheye_to_world(cam_to_pose: np.ndarray, pose_to_world: np.ndarray) -> np.ndarray:
    """image_02 / image_03 are not rectified: camera -> world = poses.txt[frame] @ calib_cam_to_pose[image_0x]."""
    return np.asarray(pose_to_world) @ np.asarray(cam_to_pose)
