"""The reference's own behaviour pins for this path are convergence assertions on synthetic scenes
(applications/badslam/src/badslam/test/test_*_optimization_*.cc).  They are re-expressed
here on the CPU oracle at reduced image size (the CUDA path runs the same checks in test_gpu_parity.py)."""
import copy

import numpy as np
import pytest

from badslam_b200 import scene as S
from oracle import cpu_oracle as O


def _scene(**kw):
    cfg = S.SceneConfig(width=160, height=120, num_keyframes=3, num_surfels=12000, cell=1, seed=11,
                        pose_noise_t=0.0, pose_noise_r=0.0, name="unit")
    for k, v in kw.items():
        setattr(cfg, k, v)
    return S.make_scene(cfg)


def test_pose_optimization_with_geometric_residual():
    """test_pose_optimization_geometric_residual.cc:50-174: start +-5 mm / +-1 mrad off, depth residuals only,
    EstimateFramePose must return to the ground truth (the reference asserts 1.1e-6 on its noise-free planes at
    640x480 with surfels created from the frame itself; our surfels come from OTHER keyframes' quantised depth, so the
    floor is the depth quantisation)."""
    sc = _scene()
    orc = O.Oracle(sc, use_depth=True, use_descriptor=False, poses=sc.poses_true)
    k = 1
    base, its0, _ = orc.estimate_frame_pose(k, sc.poses_true[k])
    offsets = [np.zeros(6)]
    for j in range(3):
        for sgn in (1, -1):
            d = np.zeros(6)
            d[j] = sgn * 0.005
            offsets.append(d)
            d = np.zeros(6)
            d[3 + j] = sgn * 0.001
            offsets.append(d)
    for d in offsets:
        start = O.se3_mul(sc.poses_true[k], O.se3_exp(d.astype(np.float32)))
        est, its, conv = orc.estimate_frame_pose(k, start)
        assert conv and its <= 30
        dt, dr = S.pose_error(est, base)
        assert dt < 2e-5 and dr < 2e-5, (d, dt, dr)     # every start converges to the same optimum
        dt, dr = S.pose_error(est, sc.poses_true[k])
        assert dt < 1.5e-3 and dr < 1e-3, (d, dt, dr)   # which is the ground truth up to depth quantisation (1 mm)


def test_pose_optimization_color_only_cues():
    """test_pose_optimization_photometric_residual.cc:50-181: descriptor residuals only; starts +-0.5 mm / +-1 mrad."""
    sc = _scene(width=320, height=240, num_surfels=20000, cell=2)
    orc = O.Oracle(sc, use_depth=False, use_descriptor=True, poses=sc.poses_true)
    k = 0
    base, _, _ = orc.estimate_frame_pose(k, sc.poses_true[k])
    for j in range(6):
        for sgn in (1, -1):
            d = np.zeros(6, np.float32)
            d[j] = sgn * (0.0005 if j < 3 else 0.001)
            est, its, conv = orc.estimate_frame_pose(k, O.se3_mul(sc.poses_true[k], O.se3_exp(d)))
            dt, dr = S.pose_error(est, base)
            assert dt < 3e-4 and dr < 2e-4, (j, sgn, dt, dr)


def test_geometry_optimization_with_geometric_residual():
    """test_geometry_optimization_geometric_residual.cc:50-222: surfel depth perturbed by <= 5 mm, poses fixed,
    10 x BundleAdjustment(geometry only) must move every observed surfel back onto the surface."""
    sc = _scene(surfel_depth_noise=0.002)
    orc = O.Oracle(sc, use_depth=True, use_descriptor=False, poses=sc.poses_true)
    n = sc.num_surfels

    def surface_distance(xyz):   # distance to the nearest scene plane (the surfel's own one, up to corners)
        pl = sc.planes.astype(np.float64)
        return np.min(np.abs(pl[:, :3] @ xyz.astype(np.float64) + pl[:, 3:4]), axis=0)

    before = surface_distance(sc.surfels[:3, :n])
    for _ in range(3):
        r = orc.bundle_adjust(optimize_poses=False, optimize_geometry=True, min_iterations=10, max_iterations=10)
        assert r.iterations_done == 10 and r.converged   # poses off => "converged" as soon as min_iterations ran (:693-701)
    after = surface_distance(orc.surfels[:3, :n])
    act = orc.active[:n] == 1
    assert act.mean() > 0.9
    assert np.median(after[act]) < 0.35 * np.median(before[act]), (np.median(after[act]), np.median(before[act]))
    assert np.percentile(after[act], 90) < 1.5e-3


def test_pcg_geometry_optimization_with_geometric_residual():
    """PCGGeometryOptimizationWithGeometricResidual (test_geometry_optimization_geometric_residual.cc:220-222): the same scene
    through BundleAdjustmentPCG (one unknown per surfel, the offset along its normal)."""
    sc = _scene(surfel_depth_noise=0.002)
    orc = O.Oracle(sc, use_depth=True, use_descriptor=False, poses=sc.poses_true)
    n = sc.num_surfels
    pl = sc.planes.astype(np.float64)
    surface_distance = lambda xyz: np.min(np.abs(pl[:, :3] @ xyz.astype(np.float64) + pl[:, 3:4]), axis=0)
    before = surface_distance(sc.surfels[:3, :n])
    for _ in range(3):
        r = orc.bundle_adjust_pcg(False, True, False, False, 10, 10, 30, 0, end_tasks=False)
        assert r.iterations_done == 10 and r.converged
    after = surface_distance(orc.surfels[:3, :n])
    assert np.median(after) < 0.35 * np.median(before) and np.percentile(after, 90) < 1.5e-3


def test_geometry_optimization_with_photometric_residual():
    """test_geometry_optimization_photometric_residual.cc:120-285: descriptors are re-estimated jointly with the
    position; after a few iterations the descriptor cost must have dropped."""
    sc = _scene(width=320, height=240, num_surfels=20000, cell=2, surfel_depth_noise=0.001)
    orc = O.Oracle(sc, poses=sc.poses_true)
    c0 = sum(orc.pose_coeffs(k).cost_desc1 + orc.pose_coeffs(k).cost_desc2 for k in range(3))
    for _ in range(5):
        orc.bundle_adjust(optimize_poses=False, optimize_geometry=True, min_iterations=1, max_iterations=1)
    c1 = sum(orc.pose_coeffs(k).cost_desc1 + orc.pose_coeffs(k).cost_desc2 for k in range(3))
    assert c1 < 0.7 * c0
    assert np.all(np.abs(orc.surfels[6:8, :sc.num_surfels]) <= 180.0)   # descriptor clamp (kernel_opt_geometry.cu:350-358)


def test_bundle_adjustment_state_machine(tiny_scene):
    """direct_ba_alternating.cc:543-577,693-717: keyframes that stop moving become inactive; BA reports convergence
    once all are; activation of surfels follows the active keyframes."""
    orc = O.Oracle(tiny_scene)

    def total_cost():
        c = 0.0
        for k in range(orc.K):
            st = orc.pose_coeffs(k)
            c += st.cost_depth + st.cost_desc1 + st.cost_desc2
        return c

    c0 = total_cost()
    r = orc.bundle_adjust(True, True, 1, 30, end_tasks=False)   # (the end tasks shrink radii, which changes the descriptor samples)
    assert r.converged and 1 < r.iterations_done <= 30
    assert np.all(orc.activation == 2)      # all inactive
    assert total_cost() < 0.8 * c0          # joint optimisation of poses and geometry lowers the cost
    # a fixed window re-activates exactly the window (+ co-visible frames)
    r = orc.bundle_adjust(True, True, 1, 1, window_start=1, window_end=1)
    assert r.iterations_done == 1


def test_empty_and_degenerate_inputs(tiny_scene):
    sc = copy.copy(tiny_scene)
    sc.num_surfels = 0
    orc = O.Oracle(sc)
    st = orc.pose_coeffs(0)
    assert st.n_assoc == 0 and not any(st.H[:])
    est, its, conv = orc.estimate_frame_pose(0)
    assert np.allclose(est, sc.poses_init[0]) and conv and its == 1     # H = 0 -> x = 0 -> converged at once
    r = orc.bundle_adjust(True, True, 1, 3)
    assert r.converged
    # surfels behind every camera associate with nothing
    sc2 = copy.copy(tiny_scene)
    sc2.surfels = tiny_scene.surfels.copy()
    sc2.surfels[2] = -5.0
    orc2 = O.Oracle(sc2)
    assert orc2.pose_coeffs(0).n_inimg == 0
    orc2.update_activation()
    assert not orc2.active.any()


# ---- the intrinsics / depth-deformation tests (test_intrinsics_optimization_*.cc), at a quarter of the reference's image size.
# reference_test_scene / empty_map_oracle are shared with tests/test_gpu_reference_tests.py, which runs the same three tests on
# the CUDA path and compares it with the oracle's result.

def reference_test_scene(seed, **kw):
    """12 keyframes looking at 20 random planes from poses spread like the reference tests' (scene.py follows
    test_intrinsics_optimization_photometric_residual.cc:182-211), exact poses, an empty surfel map with room to grow."""
    import dataclasses
    cfg = dataclasses.replace(S.SceneConfig(160, 120, 12, 2000, cell=2, seed=seed, name=f"reftest{seed}"),
                              pose_noise_t=0.0, pose_noise_r=0.0, **kw)
    sc = S.make_scene(cfg)
    sc.surfels = np.zeros((17, 1 << 17), np.float32)
    sc.num_surfels = 0
    return sc


def empty_map_oracle(sc, **kw):
    orc = O.Oracle(sc, poses=sc.poses_true, **kw)
    orc.n = 0
    return orc


DEPTH_CAMERA_PERTURBATION = 0.25 * np.array([0.5, -0.6, 1.23, -2.17])   # the reference's offsets (:150, :430) at quarter resolution


def test_depth_deformation_optimization_with_geometric_residual():
    """AlternatingDepthDeformationOptimizationWithGeometricResidual (test_intrinsics_optimization_geometric_residual.cc:178-366):
    raw depth distorted with a = 0.03, cfactor = 0.005; 400 x BundleAdjustment(max 10 iterations, surfel updates on, poses fixed,
    depth intrinsics on from the second call) starting from an EMPTY map -- the whole loop: creation, merging, deletion,
    compaction, geometry, intrinsics + deformation Schur step.  Same assertions as the reference (:347-349)."""
    sc = reference_test_scene(21, depth_a=0.03, cfactor=0.005)
    orc = empty_map_oracle(sc, use_descriptor=False)
    for i in range(400):
        orc.bundle_adjust(False, True, 1, 10, optimize_depth_intrinsics=(i != 0), do_surfel_updates=True, end_tasks=(i != 0))
    assert orc.n > 10000
    assert abs(orc.model.a - 0.03) < 1e-2
    assert abs(orc.cfactor[25, 25] - 0.005) < 1e-3
    seen = orc.cfactor != 0
    assert seen.mean() > 0.9 and abs(np.median(orc.cfactor[seen]) - 0.005) < 1e-3


def test_intrinsics_optimization_with_geometric_residual():
    """AlternatingIntrinsicsOptimizationWithGeometricResidual (:371-559): surfels created with the true camera, then the depth
    camera estimate is perturbed; 100 x BundleAdjustment(depth intrinsics only) recover it to 1e-3 px (:539-542)."""
    sc = reference_test_scene(22)
    orc = empty_map_oracle(sc, use_descriptor=False)
    for k in range(sc.cfg.num_keyframes):
        orc.create_surfels_for_keyframe(k, True)
    true_K = np.array(orc.model.depth_K[:], np.float64)
    for i in range(4):
        orc.model.depth_K[i] = float(true_K[i] + DEPTH_CAMERA_PERTURBATION[i])
    for i in range(100):
        orc.bundle_adjust(False, False, 1, 10, optimize_depth_intrinsics=True, end_tasks=(i != 0))
    assert np.all(np.abs(np.array(orc.model.depth_K[:]) - true_K) < 1e-3)


def test_intrinsics_optimization_with_photometric_residual():
    """AlternatingIntrinsicsOptimizationWithPhotometricResidual (test_intrinsics_optimization_photometric_residual.cc:105-282):
    descriptor residuals only, colour camera perturbed, 10 x BundleAdjustment(colour intrinsics only, surfel updates on);
    thresholds 0.03 / 0.03 / 0.15 / 0.15 px (:262-265)."""
    sc = reference_test_scene(23)
    orc = empty_map_oracle(sc, use_depth=False)
    for k in range(sc.cfg.num_keyframes):
        orc.create_surfels_for_keyframe(k, True)
    true_K = np.array(orc.model.color_K[:], np.float64)
    for i in range(4):
        orc.model.color_K[i] = float(true_K[i] + DEPTH_CAMERA_PERTURBATION[i])
    for i in range(10):
        orc.bundle_adjust(False, False, 1, 10, optimize_color_intrinsics=True, do_surfel_updates=True, end_tasks=(i != 0))
    err = np.abs(np.array(orc.model.color_K[:]) - true_K)
    assert np.all(err < [0.03, 0.03, 0.15, 0.15])


def test_pcg_depth_deformation_optimization_with_geometric_residual():
    """PCGDepthDeformationOptimizationWithGeometricResidual (test_intrinsics_optimization_geometric_residual.cc:364-366): the same
    test through BundleAdjustmentPCG with its surfel-update branches (direct_ba_pcg.cc:180-206,644-690,775-815); the joint
    solve needs 20 calls instead of 400."""
    sc = reference_test_scene(21, depth_a=0.03, cfactor=0.005)
    orc = empty_map_oracle(sc, use_descriptor=False)
    for i in range(20):
        r = orc.bundle_adjust_pcg(False, True, i != 0, False, 1, 10, 30, 0, end_tasks=(i != 0), do_surfel_updates=True)
        assert r.surfels_size == orc.n or i != 0          # with the end tasks the map shrinks once more after the call
    assert orc.n > 10000
    assert abs(orc.model.a - 0.03) < 1e-2
    assert abs(orc.cfactor[25, 25] - 0.005) < 1e-3


def test_pcg_intrinsics_optimization_with_geometric_and_photometric_residuals():
    """PCGIntrinsicsOptimizationWithGeometricResidual (test_intrinsics_optimization_geometric_residual.cc:565-567, 100 calls, 1e-3 px)
    and PCGIntrinsicsOptimizationWithPhotometricResidual (test_intrinsics_optimization_photometric_residual.cc:288-290)."""
    sc = reference_test_scene(22)
    orc = empty_map_oracle(sc, use_descriptor=False)
    for k in range(sc.cfg.num_keyframes):
        orc.create_surfels_for_keyframe(k, True)
    true_K = np.array(orc.model.depth_K[:], np.float64)
    for i in range(4):
        orc.model.depth_K[i] = float(true_K[i] + DEPTH_CAMERA_PERTURBATION[i])
    for i in range(100):
        orc.bundle_adjust_pcg(False, False, True, False, 1, 10, 30, 0, end_tasks=(i != 0))
    assert np.all(np.abs(np.array(orc.model.depth_K[:]) - true_K) < 1e-3)

    sc = reference_test_scene(23)
    orc = empty_map_oracle(sc, use_depth=False)
    for k in range(sc.cfg.num_keyframes):
        orc.create_surfels_for_keyframe(k, True)
    true_K = np.array(orc.model.color_K[:], np.float64)
    for i in range(4):
        orc.model.color_K[i] = float(true_K[i] + DEPTH_CAMERA_PERTURBATION[i])
    for i in range(10):
        orc.bundle_adjust_pcg(False, False, False, True, 1, 10, 30, 0, end_tasks=(i != 0), do_surfel_updates=True)
    assert np.all(np.abs(np.array(orc.model.color_K[:]) - true_K) < [0.03, 0.03, 0.15, 0.15])
