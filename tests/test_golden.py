"""CPU-only: the oracle against golden fixtures produced by the REFERENCE's own CUDA kernels (tools/make_golden.py,
run on a B200 through oracle/_ref).  This is what pins the oracle when /root/reference is not available."""
import os

import numpy as np
import pytest

from badslam_b200 import scene as S
from oracle import cpu_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


@pytest.mark.parametrize("name,tag,use_depth,use_desc", [("cfg1", "", True, True), ("tiny", "", True, True),
                                                          ("tiny", "_depth_only", True, False), ("tiny", "_desc_only", False, True)])
def test_oracle_matches_reference_cuda_golden(name, tag, use_depth, use_desc):
    path = os.path.join(GOLDEN, f"{name}{tag}.npz")
    assert os.path.exists(path), "golden fixtures are committed under tests/golden"
    g = np.load(path)
    sc = S.make_scene(S.config_by_name(name))
    # the generator is deterministic: same inputs as when the fixture was made
    assert abs(float(np.sum(sc.surfels[:3, :sc.num_surfels].astype(np.float64))) - float(g["surfel_checksum"])) < 1e-6
    assert int(sc.depth.astype(np.uint64).sum()) == int(g["depth_checksum"])
    orc = O.Oracle(sc, use_depth, use_desc)
    K = sc.cfg.num_keyframes
    for k in range(K):
        st = orc.pose_coeffs(k)
        assert (st.n_assoc if use_depth else 0) + (st.n_photo if use_desc else 0) == g["pose_count"][k]
        assert rel(st.H[:], g["pose_H"][k]) < 1e-4
        # The oracle reproduces the texture unit's bilinear filter bit-exactly (tools/tex_probe*.cu); what is left is
        # fp32 summation order and -use_fast_math.
        tol_b = 5e-4
        assert rel(st.b[:], g["pose_b"][k]) < tol_b
        cost = (st.cost_depth if use_depth else 0.0) + (st.cost_desc1 if use_desc else 0.0)
        assert abs(cost - g["pose_cost"][k]) < 2 * tol_b * g["pose_cost"][k]
        p, it, conv = orc.estimate_frame_pose(k)
        dt, dr = S.pose_error(p, g["efp_pose"][k])
        assert dt < 1e-5 and dr < 1e-5, (k, dt, dr)
        assert it == g["efp_iterations"][k] and conv == bool(g["efp_converged"][k])
    orc.update_activation()
    assert np.array_equal(np.packbits(orc.active[:sc.num_surfels]), g["activation_flags"])
    orc.optimize_geometry_iteration()
    rows = orc.surfels[[0, 1, 2, 3, 6, 7], :sc.num_surfels]
    gr = g["geometry_rows"]
    # (photometric-only position updates are ill-conditioned for low-texture surfels: only the mean is pinned there)
    assert np.max(np.abs(rows[:3] - gr[:3])) < (1e-3 if use_depth else 5e-2) and np.mean(np.abs(rows[:3] - gr[:3])) < (2e-6 if use_depth else 2e-5)
    assert (rows[3].view(np.uint32) != gr[3].view(np.uint32)).mean() < 2e-3
    assert np.mean(np.abs(rows[4:6] - gr[4:6])) < 5e-3


@pytest.mark.parametrize("name,tag,use_depth,use_desc", [("tiny", "", True, True), ("tiny", "_depth_only", True, False)])
def test_oracle_bundle_adjustment_matches_reference_cuda_golden(name, tag, use_depth, use_desc):
    g = np.load(os.path.join(GOLDEN, f"{name}{tag}.npz"))
    sc = S.make_scene(S.config_by_name(name))
    orc = O.Oracle(sc, use_depth, use_desc)
    r = orc.bundle_adjust(True, True, 3, 3, end_tasks=False)
    assert r.pose_iterations_total == int(g["ba_pose_iterations"])
    pairs = (r.n_assoc if use_depth else 0) + (r.n_photo if use_desc else 0)
    assert abs(pairs - int(g["ba_count"])) <= max(2, 1e-4 * int(g["ba_count"]))
    noise = max(max(S.pose_error(g["ba_poses"][k], g["ba_poses_rerun"][k])) for k in range(orc.K))
    for k in range(orc.K):
        dt, dr = S.pose_error(orc.poses[k], g["ba_poses"][k])
        assert dt < 5e-5 + 2 * noise and dr < 5e-5 + 2 * noise, (k, dt, dr, noise)
    assert np.array_equal(orc.activation, g["ba_activation"])
