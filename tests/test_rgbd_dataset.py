"""CPU-only: the TUM / ETH3D "associated and calibrated" RGB-D dataset layout the reference reads
(libvis/src/libvis/rgbd_video_io_tum_dataset.h:42-232)."""
import os

import numpy as np
import pytest

from badslam_b200 import rgbd_dataset as D
from badslam_b200 import scene as S


def test_write_then_read_gives_the_same_frames(tmp_path, tiny_scene):
    sc = tiny_scene
    K = sc.cfg.num_keyframes
    frames = [S.raw_frame(sc, k) for k in range(K)]
    stamps = [1305031102.175304 + 0.033 * k for k in range(K)]
    D.write_tum_dataset(str(tmp_path), sc.depth_K, [f[1] for f in frames], [f[0] for f in frames], stamps, sc.poses_true)
    ds = D.TUMRGBDDataset(str(tmp_path), "groundtruth.txt")
    assert len(ds) == K and (ds.width, ds.height) == (sc.cfg.width, sc.cfg.height)
    assert np.allclose(ds.camera_parameters, sc.depth_K, atol=1e-4)          # +0.5 on load undoes the -0.5 on disk
    for k in range(K):
        assert np.array_equal(ds.load_depth(k), frames[k][0])                # 16-bit PNG: lossless
        assert np.array_equal(ds.load_color(k), frames[k][1])
        fr = ds.frames[k]
        assert abs(fr.rgb_timestamp - stamps[k]) < 1e-6 and fr.rgb_time_string == f"{stamps[k]:.6f}"
        assert S.pose_error(fr.depth_global_T_frame, sc.poses_true[k])[0] < 1e-6
        assert S.pose_error(fr.rgb_global_T_frame, sc.poses_true[k])[1] < 1e-6
    without = D.TUMRGBDDataset(str(tmp_path))
    assert len(without) == K and without.frames[0].rgb_global_T_frame is None


def test_trajectory_parsing_and_pose_interpolation(tmp_path):
    p = tmp_path / "traj.txt"
    half = np.sin(np.pi / 4)        # 90 degrees about z
    p.write_text("# ground truth trajectory\n# timestamp tx ty tz qx qy qz qw\n"
                 "10.0 0 0 0 0 0 0 1\n"
                 f"12.0 2 4 6 0 0 {half} {half}\n"
                 "\n"
                 "99.0 9 9 9 0 0 0 1\n")          # after the first empty line: not read (the reference stops there)
    ts, poses = D.read_tum_trajectory(str(p))
    assert list(ts) == [10.0, 12.0] and poses.shape == (2, 7)
    assert np.allclose(poses[1], [0, 0, half, half, 2, 4, 6])               # stored as qx qy qz qw tx ty tz
    assert np.allclose(D.interpolate_pose(5.0, ts, poses), poses[0])         # clamped before the first pose
    assert np.allclose(D.interpolate_pose(50.0, ts, poses), poses[1])        # and after the last
    mid = D.interpolate_pose(11.0, ts, poses)
    q45 = [0, 0, np.sin(np.pi / 8), np.cos(np.pi / 8)]
    assert np.allclose(mid[:4], q45, atol=1e-6) and np.allclose(mid[4:], [1, 2, 3])
    # the shortest arc is taken when the quaternions have opposite signs
    flipped = poses.copy()
    flipped[1, :4] *= -1
    mid2 = D.interpolate_pose(11.0, ts, flipped)
    assert np.allclose(np.abs(mid2[:4]), np.abs(q45), atol=1e-6) and abs(np.linalg.norm(mid2[:4]) - 1) < 1e-6


def test_malformed_inputs_raise(tmp_path):
    with pytest.raises(OSError):
        D.TUMRGBDDataset(str(tmp_path))                       # no calibration.txt
    (tmp_path / "calibration.txt").write_text("525 525 319.5\n")
    with pytest.raises(ValueError):
        D.TUMRGBDDataset(str(tmp_path))
    (tmp_path / "calibration.txt").write_text("525 525 319.5 239.5\n")
    (tmp_path / "associated.txt").write_text("# nothing\n")
    with pytest.raises(ValueError):
        D.TUMRGBDDataset(str(tmp_path))
    (tmp_path / "associated.txt").write_text("1.0 rgb/1.png 1.0\n")
    with pytest.raises(ValueError):
        D.TUMRGBDDataset(str(tmp_path))
    (tmp_path / "associated.txt").write_text("1.0 rgb/1.png 1.0 depth/1.png\n")
    with pytest.raises(OSError):
        D.TUMRGBDDataset(str(tmp_path))                       # the image files are missing


def test_saved_poses_read_back_relative_to_the_start_frame(tmp_path, tiny_scene):
    """SavePoses (io.cc:537-568) writes what ReadTUMRGBDTrajectory reads; the start frame ends up at the identity."""
    sc = tiny_scene
    K = sc.cfg.num_keyframes
    names = [f"{100.0 + k:.6f}" for k in range(K)]
    path = str(tmp_path / "poses.txt")
    assert D.save_poses(path, names, sc.poses_true, start_frame=1)
    ts, poses = D.read_tum_trajectory(path)
    assert list(ts) == [100.0 + k for k in range(K)]
    assert S.pose_error(poses[1], [0, 0, 0, 1, 0, 0, 0]) < (1e-6, 1e-6)
    for k in range(K):
        want = S.se3_mul(S.se3_inverse(sc.poses_true[1]), sc.poses_true[k])
        e = S.pose_error(poses[k], want)
        assert e[0] < 1e-5 and e[1] < 1e-5
    assert open(path).readline().startswith("# Format: Each line gives one global_T_frame pose")
    assert not D.save_poses(str(tmp_path / "no_such_dir" / "poses.txt"), names, sc.poses_true)
