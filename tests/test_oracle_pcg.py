"""CPU-only: the oracle's intrinsics step and PCG solver (a) against golden vectors produced by the REFERENCE's own CUDA
kernels (tools/make_golden.py::golden_intrinsics_pcg, run on a B200 through oracle/_ref) and (b) through properties."""
import dataclasses
import os

import numpy as np
import pytest

from badslam_b200 import scene as S
from oracle import cpu_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_intrinsics_pcg.npz")


def distorted_scene(name="tiny"):
    """Must stay identical to tools/make_golden.py::distorted_scene."""
    sc = S.make_scene(dataclasses.replace(S.config_by_name(name), depth_a=0.03, cfactor=0.005))
    sc.depth_K = (np.asarray(sc.depth_K, np.float32) * np.float32([1.003, 0.998, 1.002, 0.997])).astype(np.float32)
    sc.color_K = (np.asarray(sc.color_K, np.float32) * np.float32([0.998, 1.002, 1.001, 0.999])).astype(np.float32)
    a_init = 0.02
    cf_init = (np.random.default_rng(5).standard_normal(sc.cfactor.shape) * 0.003).astype(np.float32)
    return sc, a_init, cf_init


def make_oracle(sc, a_init, cf_init):
    orc = O.Oracle(sc)
    orc.model.a = a_init
    orc.cfactor[:] = cf_init
    return orc


@pytest.fixture(scope="module")
def setup():
    g = np.load(GOLDEN)
    sc, a_init, cf_init = distorted_scene("tiny")
    assert abs(float(np.sum(sc.surfels[:3, :sc.num_surfels].astype(np.float64))) - float(g["surfel_checksum"])) < 1e-6
    return g, sc, a_init, cf_init


def test_intrinsics_step_matches_reference_cuda_golden(setup):
    """OptimizeIntrinsicsCUDA, kernel_opt_intrinsics.cc:39-281 (two consecutive steps, non-zero a / cfactor)."""
    g, sc, a_init, cf_init = setup
    orc = make_oracle(sc, a_init, cf_init)
    for step in range(2):
        orc.optimize_intrinsics(True, True)
        d, c = np.array(orc.model.depth_K[:], np.float32), np.array(orc.model.color_K[:], np.float32)
        assert np.abs(d - g[f"intr{step}_depth_K"]).max() < 2e-3, (step, d, g[f"intr{step}_depth_K"])
        assert np.abs(c - g[f"intr{step}_color_K"]).max() < 2e-3
        assert abs(orc.model.a - float(g[f"intr{step}_a"])) < 2e-5
        assert np.abs(orc.cfactor - g[f"intr{step}_cfactor"]).max() < 1e-4
    assert abs(orc.model.a - a_init) > 1e-3 and np.abs(np.array(orc.model.depth_K[:]) - sc.depth_K).max() > 0.1


@pytest.mark.parametrize("intr", [False, True])
def test_pcg_building_blocks_match_reference_cuda_golden(setup, intr):
    """r, M after PCGInit; p0; g = J^T W J p0; alpha_n, alpha_d (kernel_pcg.cu:179-1037)."""
    g, sc, a_init, cf_init = setup
    orc = make_oracle(sc, a_init, cf_init)
    K, n = sc.cfg.num_keyframes, sc.num_surfels
    r, M, p, gv, scal = orc.pcg_debug(optimize_depth_intrinsics=intr, optimize_color_intrinsics=intr, gauge_keyframe=1)
    tag = "pcgi" if intr else "pcg"
    lo, hi = 6 * (K - 1), 6 * (K - 1) + 3 * n
    assert len(r) == hi + ((5 + sc.cfactor.size + 4) if intr else 0)
    for nm, v in (("r", r), ("M", M), ("p", p), ("g", gv)):
        ref = g[f"{tag}_{nm}_pose"]
        assert np.abs(v[:lo] - ref).max() < 5e-4 * np.abs(ref).max(), nm
        # surfel unknowns: aggregate only (single entries flip with association thresholds / the texture filter emulation)
        assert abs(v[lo:hi].astype(np.float64).sum() - float(g[f"{tag}_{nm}_surfel_sum"])) < 2e-3 * float(g[f"{tag}_{nm}_surfel_abs"]), nm
        if intr:
            ref = g[f"{tag}_{nm}_intr"]
            glob = np.r_[0:5, len(ref) - 4:len(ref)]      # fx^-1 fy^-1 cx^-1 cy^-1 a ... colour fx fy cx cy
            assert np.abs(v[hi:][glob] - ref[glob]).max() < 5e-4 * np.abs(ref[glob]).max(), nm
            assert np.abs(v[hi:] - ref).max() < 2e-2 * np.abs(ref).max(), nm      # per-cell cfactor entries
    assert np.all(np.abs(scal - g[f"{tag}_scalars"]) < 2e-4 * np.abs(g[f"{tag}_scalars"]))


@pytest.mark.parametrize("intr", [False, True])
def test_pcg_short_solve_matches_reference_cuda_golden(setup, intr):
    """Two outer iterations x 4 inner PCG steps (few steps: the fp32 CG recurrences have not decorrelated yet)."""
    g, sc, a_init, cf_init = setup
    orc = make_oracle(sc, a_init, cf_init)
    res = orc.bundle_adjust_pcg(True, True, intr, intr, 2, 2, 4, 1, end_tasks=False)
    tag = "pcgi" if intr else "pcg"
    assert res.iterations_done == 2 and res.inner_iterations_total == 8
    noise = max(max(S.pose_error(g[f"{tag}_ba_poses"][k], g[f"{tag}_ba_poses_rerun"][k])) for k in range(orc.K))
    for k in range(orc.K):
        dt, dr = S.pose_error(orc.poses[k], g[f"{tag}_ba_poses"][k])
        assert dt < 5e-5 + 3 * noise and dr < 5e-5 + 3 * noise, (k, dt, dr, noise)
    assert abs(res.last_r_norm - float(g[f"{tag}_ba_r_norm"])) < 2e-2 * float(g[f"{tag}_ba_r_norm"])
    assert np.mean(np.abs(orc.surfels[:3, :sc.num_surfels] - g[f"{tag}_ba_surfels"][:3])) < 1e-5
    if intr:
        assert np.abs(np.array(orc.model.depth_K[:]) - g["pcgi_ba_depth_K"]).max() < 5e-3
        assert np.abs(np.array(orc.model.color_K[:]) - g["pcgi_ba_color_K"]).max() < 5e-3
        assert abs(orc.model.a - float(g["pcgi_ba_a"])) < 1e-3


def test_pcg_converges_and_keeps_the_gauge_keyframe():
    sc = S.make_scene(S.config_by_name("tiny"))
    orc = O.Oracle(sc)
    K = sc.cfg.num_keyframes

    def rel_err(poses):
        e = []
        for k in range(1, K):
            a = O.se3_mul(O.se3_inverse(poses[0]), poses[k])
            b = O.se3_mul(O.se3_inverse(sc.poses_true[0]), sc.poses_true[k])
            e.append(max(S.pose_error(a, b)))
        return max(e)

    e0 = rel_err(orc.poses)
    res = orc.bundle_adjust_pcg(min_iterations=5, max_iterations=5, gauge_keyframe=0, end_tasks=False)
    assert res.iterations_done == 5 and 5 <= res.inner_iterations_total <= 150
    assert np.array_equal(orc.poses[0], sc.poses_init[0])
    assert rel_err(orc.poses) < 0.2 * e0


def test_pcg_preconditioner_and_first_step_are_consistent():
    """alpha_n = r^T p, alpha_d = p^T g + K * lambda |p|^2 (the reference adds the lambda term once per keyframe launch,
    kernel_pcg.cu:1101-1112), M = diag(J^T W J) >= 0 and g = A p is a descent-consistent direction (p^T g > 0)."""
    sc = S.make_scene(S.config_by_name("tiny"))
    orc = O.Oracle(sc)
    r, M, p, g, scal = orc.pcg_debug(gauge_keyframe=2)
    assert np.all(M >= 0)
    assert abs(float(np.dot(r.astype(np.float64), p)) - scal[0]) < 1e-6 * scal[0]
    lam = 1e-8
    expect = float(np.dot(p.astype(np.float64), g)) + sc.cfg.num_keyframes * lam * float(np.dot(p.astype(np.float64), p))
    assert abs(expect - scal[1]) < 1e-5 * scal[1]
    np.testing.assert_allclose(p, r / (M + np.float32(lam)), rtol=1e-6)
