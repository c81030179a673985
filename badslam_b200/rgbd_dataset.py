"""Reader (and writer) for the RGB-D dataset layout the reference consumes: the "associated and calibrated" variant of the TUM
RGB-D format that the ETH3D SLAM benchmark ships (ReadTUMRGBDDatasetAssociatedAndCalibrated,
libvis/src/libvis/rgbd_video_io_tum_dataset.h:112-232):

    <folder>/calibration.txt      "fx fy cx cy" on one line, pixel-CENTRE convention (0.5 is added to cx, cy on load, :221-224)
    <folder>/associated.txt       one line per frame: "<rgb time> <rgb file> <depth time> <depth file>", '#' starts a comment
    <folder>/rgb/*.png            8-bit colour
    <folder>/depth/*.png          16-bit raw depth (0 = no measurement; 5000 units per metre in TUM / ETH3D data)
    <folder>/<trajectory>         optional, TUM trajectory lines "time tx ty tz qx qy qz qw" (ReadTUMRGBDTrajectory, :68-110);
                                  frame poses are interpolated at the image time stamps (InterpolatePose, :42-66)

This is the input side of the hot path: frames from here go through DirectBA.PreprocessFrame (bba_preprocess_frame) and
AddKeyframe.  Poses use the library's layout {qx, qy, qz, qw, tx, ty, tz} (global_T_frame).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class RGBDFrame:
    rgb_timestamp: float
    rgb_time_string: str
    rgb_path: str
    depth_timestamp: float
    depth_time_string: str
    depth_path: str
    rgb_global_T_frame: Optional[np.ndarray] = None      # [7], None without a trajectory
    depth_global_T_frame: Optional[np.ndarray] = None


def read_tum_trajectory(path: str):
    """ReadTUMRGBDTrajectory (:68-110): returns (timestamps [N] float64, poses [N, 7] float32 as qx qy qz qw tx ty tz).  Reading
    stops at the first empty line like the reference does; lines starting with '#' are skipped."""
    stamps, poses = [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line == "":
                break
            if line[0] == "#":
                continue
            tok = line.split()
            if len(tok) < 8:
                raise ValueError(f"cannot read pose line: {line!r}")
            t = [float(v) for v in tok[1:8]]
            stamps.append(float(tok[0]))
            poses.append([t[3], t[4], t[5], t[6], t[0], t[1], t[2]])
    return np.asarray(stamps, np.float64), np.asarray(poses, np.float32).reshape(-1, 7)


def _slerp(qa, qb, t):
    """Eigen::Quaternion::slerp (the reference interpolates with it, :59): shortest arc, linear for nearly parallel inputs."""
    qa, qb = np.asarray(qa, np.float64), np.asarray(qb, np.float64)
    d = float(np.dot(qa, qb))
    ad = abs(d)
    if ad >= 1.0 - np.finfo(np.float32).eps:
        s0, s1 = 1.0 - t, t
    else:
        theta = np.arccos(ad)
        st = np.sin(theta)
        s0, s1 = np.sin((1.0 - t) * theta) / st, np.sin(t * theta) / st
    if d < 0:
        s1 = -s1
    return s0 * qa + s1 * qb


def interpolate_pose(timestamp: float, pose_timestamps, poses):
    """InterpolatePose (:42-66): clamps outside the trajectory, slerp + linear translation inside; None if no bracket is found."""
    ts = np.asarray(pose_timestamps, np.float64)
    assert len(ts) == len(poses) and len(ts) >= 2
    if timestamp <= ts[0]:
        return np.asarray(poses[0], np.float32).copy()
    if timestamp >= ts[-1]:
        return np.asarray(poses[-1], np.float32).copy()
    i = int(np.searchsorted(ts, timestamp, side="right")) - 1
    if i < 0 or i + 1 >= len(ts) or not (ts[i] <= timestamp <= ts[i + 1]):
        return None
    f = (timestamp - ts[i]) / (ts[i + 1] - ts[i])
    a, b = np.asarray(poses[i], np.float64), np.asarray(poses[i + 1], np.float64)
    out = np.empty(7, np.float32)
    out[:4] = _slerp(a[:4], b[:4], f)
    out[4:] = a[4:] + f * (b[4:] - a[4:])
    return out


class TUMRGBDDataset:
    """ReadTUMRGBDDatasetAssociatedAndCalibrated (:112-232).  Images are loaded lazily, like the reference's ImageFrame."""

    def __init__(self, dataset_folder_path: str, trajectory_filename: Optional[str] = None):
        self.folder = dataset_folder_path
        with open(os.path.join(self.folder, "calibration.txt")) as f:
            tok = f.readline().split()
        if len(tok) < 4:
            raise ValueError("cannot read calibration")
        fx, fy, cx, cy = [float(v) for v in tok[:4]]
        stamps, poses = None, None
        if trajectory_filename:
            stamps, poses = read_tum_trajectory(os.path.join(self.folder, trajectory_filename))
        self.frames: List[RGBDFrame] = []
        with open(os.path.join(self.folder, "associated.txt")) as f:
            for line in f:
                line = line.strip()
                if not line or line[0] == "#":
                    continue
                tok = line.split()
                if len(tok) < 4:
                    raise ValueError(f"cannot read association line: {line!r}")
                fr = RGBDFrame(float(tok[0]), tok[0], os.path.join(self.folder, tok[1]),
                               float(tok[2]), tok[2], os.path.join(self.folder, tok[3]))
                if poses is not None and len(poses):
                    fr.rgb_global_T_frame = interpolate_pose(fr.rgb_timestamp, stamps, poses)
                    fr.depth_global_T_frame = interpolate_pose(fr.depth_timestamp, stamps, poses)
                    if fr.rgb_global_T_frame is None or fr.depth_global_T_frame is None:
                        continue
                self.frames.append(fr)
        if not self.frames:
            raise ValueError("no frames in associated.txt")
        h, w = self.load_color(0).shape[:2]
        self.width, self.height = int(w), int(h)
        # PinholeCamera4f parameters in the pixel-corner convention the backend uses (:221-224); colour = depth camera (:225-228)
        self.camera_parameters = np.array([fx, fy, cx + 0.5, cy + 0.5], np.float32)

    def __len__(self):
        return len(self.frames)

    def load_color(self, i: int) -> np.ndarray:
        """[h, w, 3] uint8, RGB order (the uchar3 image bba_preprocess_frame takes)."""
        import cv2
        img = cv2.imread(self.frames[i].rgb_path, cv2.IMREAD_COLOR)
        if img is None:
            raise OSError(f"cannot load {self.frames[i].rgb_path}")
        return np.ascontiguousarray(img[..., ::-1])

    def load_depth(self, i: int) -> np.ndarray:
        """[h, w] uint16 raw depth, 0 = no measurement."""
        import cv2
        img = cv2.imread(self.frames[i].depth_path, cv2.IMREAD_UNCHANGED)
        if img is None or img.dtype != np.uint16 or img.ndim != 2:
            raise OSError(f"cannot load {self.frames[i].depth_path} as a 16-bit depth image")
        return img


def write_tum_dataset(folder: str, camera_parameters, colors, depths, timestamps, poses=None, trajectory_filename="groundtruth.txt"):
    """Writes a dataset in the layout above (tests, tools): colours [h, w, 3] RGB uint8, depths [h, w] uint16, poses [N, 7]
    qx qy qz qw tx ty tz.  camera_parameters are pixel-corner fx fy cx cy (0.5 is subtracted on disk)."""
    import cv2
    os.makedirs(os.path.join(folder, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(folder, "depth"), exist_ok=True)
    K = np.asarray(camera_parameters, np.float64)
    with open(os.path.join(folder, "calibration.txt"), "w") as f:
        f.write(f"{K[0]:.9g} {K[1]:.9g} {K[2] - 0.5:.9g} {K[3] - 0.5:.9g}\n")
    with open(os.path.join(folder, "associated.txt"), "w") as f:
        for i, t in enumerate(timestamps):
            name = f"{t:.6f}"
            cv2.imwrite(os.path.join(folder, "rgb", name + ".png"), np.ascontiguousarray(np.asarray(colors[i])[..., ::-1]))
            cv2.imwrite(os.path.join(folder, "depth", name + ".png"), np.ascontiguousarray(depths[i], dtype=np.uint16))
            f.write(f"{name} rgb/{name}.png {name} depth/{name}.png\n")
    if poses is not None:
        with open(os.path.join(folder, trajectory_filename), "w") as f:
            f.write("# timestamp tx ty tz qx qy qz qw\n")
            for t, p in zip(timestamps, poses):
                p = np.asarray(p, np.float64)
                f.write(f"{t:.6f} {p[4]:.9g} {p[5]:.9g} {p[6]:.9g} {p[0]:.9g} {p[1]:.9g} {p[2]:.9g} {p[3]:.9g}\n")


def save_poses(export_poses_path: str, time_strings, poses_global_T_frame, start_frame: int = 0) -> bool:
    """SavePoses (io.cc:537-568): the trajectory in TUM format, "time tx ty tz qx qy qz qw" with 16 significant digits, every
    pose pre-multiplied by frame_T_global of `start_frame` so that that frame sits at the identity; time stamps are written as
    the strings they were read with."""
    from . import scene as S
    P = np.asarray(poses_global_T_frame, np.float64)
    start_frame_T_global = S.se3_inverse(P[start_frame])
    try:
        f = open(export_poses_path, "w")
    except OSError:
        return False
    with f:
        f.write("# Format: Each line gives one global_T_frame pose with values: tx ty tz qx qy qz qw\n")
        for ts, p in zip(time_strings, P):
            g = np.asarray(S.se3_mul(start_frame_T_global, p), np.float32)   # SE3f arithmetic, then printed through double
            vals = [g[4], g[5], g[6], g[0], g[1], g[2], g[3]]
            f.write(str(ts) + " " + " ".join("%.16g" % float(v) for v in vals) + "\n")
    return True
