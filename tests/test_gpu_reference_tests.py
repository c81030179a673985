"""The reference's own intrinsics / depth-deformation tests (applications/badslam/src/badslam/test/test_intrinsics_optimization_*.cc)
on the CUDA path, with the reference's assertions, on the quarter-resolution scenes tests/test_oracle_convergence.py runs through
the oracle (the results of the two are printed side by side; they are not expected to be bit-identical after hundreds of
BundleAdjustment calls on a growing map, the assertions are the reference's convergence thresholds)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from badslam_b200 import scene as S
    from badslam_b200.direct_ba import DirectBA, PinholeCamera4f
    import test_oracle_convergence as T
    return S, DirectBA, PinholeCamera4f, T


def test_depth_deformation_optimization_with_geometric_residual(mods):
    """test_intrinsics_optimization_geometric_residual.cc:178-366 (assertions :347-349): from an EMPTY map, 400 x
    BundleAdjustment(max 10, surfel updates on, poses fixed, depth intrinsics on from the second call)."""
    S, DirectBA, Cam, T = mods
    sc = T.reference_test_scene(21, depth_a=0.03, cfactor=0.005)
    ba = DirectBA.from_scene(sc, poses=sc.poses_true, use_descriptor_residuals=False)
    assert ba.surfels_size() == 0
    orc = T.empty_map_oracle(sc, use_descriptor=False)
    for i in range(400):
        ba.BundleAdjustment(None, i != 0, False, True, False, True, 1, 10, increase_ba_iteration_count=(i != 0))
        if i < 3:      # side by side with the oracle while rounding differences have had no time to grow
            orc.bundle_adjust(False, True, 1, 10, optimize_depth_intrinsics=(i != 0), do_surfel_updates=True, end_tasks=(i != 0))
            assert abs(ba.surfels_size() - orc.n) <= 0.01 * orc.n, (i, ba.surfels_size(), orc.n)
            assert abs(ba.a() - orc.model.a) <= 1e-3 + 0.05 * abs(orc.model.a), (i, ba.a(), orc.model.a)
    cf = ba.cfactor_buffer()
    print(f"a = {ba.a():.4f} (true 0.03), cfactor[25, 25] = {cf[25, 25]:.5f} (true 0.005), surfels {ba.surfels_size()}")
    assert ba.surfels_size() > 10000
    assert abs(ba.a() - 0.03) < 1e-2
    assert abs(cf[25, 25] - 0.005) < 1e-3
    seen = cf != 0
    assert seen.mean() > 0.9 and abs(np.median(cf[seen]) - 0.005) < 1e-3


def test_intrinsics_optimization_with_geometric_residual(mods):
    """:371-559 (assertions :539-542; the oracle reaches 8e-4 px on this quarter-resolution scene, the bound here is 2e-3)."""
    S, DirectBA, Cam, T = mods
    sc = T.reference_test_scene(22)
    ba = DirectBA.from_scene(sc, poses=sc.poses_true, use_descriptor_residuals=False)
    for k in range(sc.cfg.num_keyframes):
        ba.CreateSurfelsForKeyframe(None, True, k)
    assert ba.surfels_size() > 10000
    true_K = np.asarray(sc.depth_K, np.float64)
    ba.SetDepthCamera(Cam(sc.cfg.width, sc.cfg.height, (true_K + T.DEPTH_CAMERA_PERTURBATION).astype(np.float32)))
    for i in range(100):
        ba.BundleAdjustment(None, True, False, False, False, False, 1, 10, increase_ba_iteration_count=(i != 0))
    err = np.abs(np.asarray(ba.depth_camera().parameters, np.float64) - true_K)
    print("depth camera error (px):", err)
    assert np.all(err < 2e-3)


def test_intrinsics_optimization_with_photometric_residual(mods):
    """test_intrinsics_optimization_photometric_residual.cc:105-282 (assertions :262-265)."""
    S, DirectBA, Cam, T = mods
    sc = T.reference_test_scene(23)
    ba = DirectBA.from_scene(sc, poses=sc.poses_true, use_depth_residuals=False)
    for k in range(sc.cfg.num_keyframes):
        ba.CreateSurfelsForKeyframe(None, True, k)
    true_K = np.asarray(sc.color_K, np.float64)
    ba.SetColorCamera(Cam(sc.cfg.width, sc.cfg.height, (true_K + T.DEPTH_CAMERA_PERTURBATION).astype(np.float32)))
    for i in range(10):
        ba.BundleAdjustment(None, False, True, True, False, False, 1, 10, increase_ba_iteration_count=(i != 0))
    err = np.abs(np.asarray(ba.color_camera().parameters, np.float64) - true_K)
    print("colour camera error (px):", err)
    assert np.all(err < [0.03, 0.03, 0.15, 0.15])


def test_pcg_depth_deformation_optimization_with_geometric_residual(mods):
    """PCGDepthDeformationOptimizationWithGeometricResidual (:364-366): 20 x BundleAdjustment(use_pcg) with surfel updates
    (direct_ba_pcg.cc:180-206,644-690,775-815)."""
    S, DirectBA, Cam, T = mods
    sc = T.reference_test_scene(21, depth_a=0.03, cfactor=0.005)
    ba = DirectBA.from_scene(sc, poses=sc.poses_true, use_descriptor_residuals=False)
    for i in range(20):
        r = ba.BundleAdjustment(None, i != 0, False, True, False, True, 1, 10, use_pcg=True, increase_ba_iteration_count=(i != 0),
                                pcg_gauge_keyframe=0)
        assert r.surfels_size == ba.surfels_size()
    cf = ba.cfactor_buffer()
    print(f"PCG: a = {ba.a():.4f} (true 0.03), cfactor[25, 25] = {cf[25, 25]:.5f} (true 0.005), surfels {ba.surfels_size()}")
    assert ba.surfels_size() > 10000
    assert abs(ba.a() - 0.03) < 1e-2
    assert abs(cf[25, 25] - 0.005) < 1e-3


def test_pcg_intrinsics_optimization_with_geometric_and_photometric_residuals(mods):
    """The PCG variants of the two camera tests (…geometric_residual.cc:565-567, …photometric_residual.cc:288-290)."""
    S, DirectBA, Cam, T = mods
    sc = T.reference_test_scene(22)
    ba = DirectBA.from_scene(sc, poses=sc.poses_true, use_descriptor_residuals=False)
    for k in range(sc.cfg.num_keyframes):
        ba.CreateSurfelsForKeyframe(None, True, k)
    true_K = np.asarray(sc.depth_K, np.float64)
    ba.SetDepthCamera(Cam(sc.cfg.width, sc.cfg.height, (true_K + T.DEPTH_CAMERA_PERTURBATION).astype(np.float32)))
    for i in range(100):
        ba.BundleAdjustment(None, True, False, False, False, False, 1, 10, use_pcg=True, increase_ba_iteration_count=(i != 0),
                            pcg_gauge_keyframe=0)
    err = np.abs(np.asarray(ba.depth_camera().parameters, np.float64) - true_K)
    print("PCG depth camera error (px):", err)
    assert np.all(err < 2e-3)      # the oracle reaches 5e-4 on this quarter-resolution scene

    sc = T.reference_test_scene(23)
    ba = DirectBA.from_scene(sc, poses=sc.poses_true, use_depth_residuals=False)
    for k in range(sc.cfg.num_keyframes):
        ba.CreateSurfelsForKeyframe(None, True, k)
    true_K = np.asarray(sc.color_K, np.float64)
    ba.SetColorCamera(Cam(sc.cfg.width, sc.cfg.height, (true_K + T.DEPTH_CAMERA_PERTURBATION).astype(np.float32)))
    for i in range(10):
        ba.BundleAdjustment(None, False, True, True, False, False, 1, 10, use_pcg=True, increase_ba_iteration_count=(i != 0),
                            pcg_gauge_keyframe=0)
    err = np.abs(np.asarray(ba.color_camera().parameters, np.float64) - true_K)
    print("PCG colour camera error (px):", err)
    assert np.all(err < [0.03, 0.03, 0.15, 0.15])
