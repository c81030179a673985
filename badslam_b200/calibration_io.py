"""Calibration text IO in the reference's file format (SaveCalibration / LoadCalibration, io.h:60-72, io.cc:570-700): the depth
and colour intrinsics and the depth deformation a DirectBA has estimated, as three text files next to each other:

    <base>.depth_intrinsics.txt   "fx fy cx-0.5 cy-0.5"       (pixel-centre convention on disk, pixel-corner in memory)
    <base>.color_intrinsics.txt   the same for the colour camera
    <base>.deformation.txt        "width height" / a / the cfactor grid row by row, one value per line (8 significant digits)

`write_calibration` / `read_calibration` work on plain arrays (no GPU); `SaveCalibration` / `LoadCalibration` apply them to a
DirectBA like the reference's functions do.
"""
from __future__ import annotations

import numpy as np


def _g(x, digits):   # what `ostream << float` prints at the given precision
    return ("%." + str(digits) + "g") % float(x)


def write_calibration(base_path: str, depth_intrinsics, color_intrinsics, a: float, cfactor: np.ndarray) -> None:
    """io.cc:570-623."""
    for suffix, K in ((".depth_intrinsics.txt", depth_intrinsics), (".color_intrinsics.txt", color_intrinsics)):
        K = np.asarray(K, np.float32)
        with open(base_path + suffix, "w") as f:   # parameters()[2] - 0.5 is evaluated in double (io.cc:582-583)
            f.write(" ".join([_g(K[0], 6), _g(K[1], 6), _g(float(K[2]) - 0.5, 6), _g(float(K[3]) - 0.5, 6)]))
    cf = np.asarray(cfactor, np.float32)
    with open(base_path + ".deformation.txt", "w") as f:
        f.write(f"{cf.shape[1]} {cf.shape[0]}\n")
        f.write(_g(a, 8) + "\n")
        for v in cf.reshape(-1):
            f.write(_g(v, 8) + "\n")


def read_calibration(base_path: str, expected_cfactor_shape=None):
    """io.cc:626-700.  Returns (depth_intrinsics[4], color_intrinsics[4], a, cfactor[h, w]); raises ValueError when the cfactor grid
    on disk has a different size than `expected_cfactor_shape` (the reference refuses that too, io.cc:676-680)."""
    out = []
    for suffix in (".depth_intrinsics.txt", ".color_intrinsics.txt"):
        with open(base_path + suffix) as f:
            vals = f.read().split()
        if len(vals) < 4:
            raise ValueError(f"{base_path + suffix}: expected 4 values")
        K = np.array([float(v) for v in vals[:4]], np.float32)
        K[2] += np.float32(0.5)
        K[3] += np.float32(0.5)
        out.append(K)
    with open(base_path + ".deformation.txt") as f:
        vals = f.read().split()
    w, h = int(vals[0]), int(vals[1])
    if expected_cfactor_shape is not None and (h, w) != tuple(expected_cfactor_shape):
        raise ValueError(f"cfactor grid {w}x{h} on disk, {expected_cfactor_shape[1]}x{expected_cfactor_shape[0]} in the current configuration")
    a = float(np.float32(vals[2]))
    if len(vals) < 3 + w * h:
        raise ValueError(f"{base_path}.deformation.txt: {len(vals) - 3} cfactor values, expected {w * h}")
    cf = np.array([float(v) for v in vals[3:3 + w * h]], np.float32).reshape(h, w)
    return out[0], out[1], a, cf


def SaveCalibration(direct_ba, export_base_path: str, stream=None) -> bool:
    write_calibration(export_base_path, direct_ba.depth_camera().parameters, direct_ba.color_camera().parameters, direct_ba.a(),
                      direct_ba.cfactor_buffer(stream))
    return True


def LoadCalibration(direct_ba, import_base_path: str, stream=None) -> bool:
    from .direct_ba import PinholeCamera4f
    try:
        d, c, a, cf = read_calibration(import_base_path, direct_ba.cfactor_buffer(stream).shape)
    except (OSError, ValueError):
        return False
    dc, cc = direct_ba.depth_camera(), direct_ba.color_camera()
    direct_ba.SetDepthCamera(PinholeCamera4f(dc.width, dc.height, d))
    direct_ba.SetColorCamera(PinholeCamera4f(cc.width, cc.height, c))
    direct_ba.SetA(a)
    direct_ba.SetCFactorBuffer(cf, stream)
    return True
