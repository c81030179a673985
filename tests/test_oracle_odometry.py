"""CPU-only: the numpy oracle of the image-pair odometry (oracle/odometry_oracle.py) -- structural properties, a numeric check
of its pose Jacobians, convergence on a rendered frame pair, and the golden fixture produced by the REFERENCE's own kernels on a
B200 (tests/golden/tiny_odometry.npz, tools/make_golden.py --odometry-only)."""
import os

import numpy as np
import pytest

from badslam_b200 import scene as S
from oracle import odometry_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_odometry.npz")
MOTION = [0.02, -0.01, 0.015, 0.01, -0.008, 0.012]
IDENT = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


@pytest.fixture(scope="module")
def pair(tiny_scene):
    sc = tiny_scene
    true_rel = S.se3_exp(MOTION)
    depth, normals, _, color = S.render_frame(sc, S.se3_mul(sc.poses_true[0], true_rel))
    return sc, true_rel, (depth, normals, color)


def make(sc, frame, num_scales=3, **kw):
    build_kw = {k: kw.pop(k) for k in ("use_pyramid_level_0", "use_gradmag") if k in kw}
    od = O.Odometry(sc.depth_K, sc.color_K, sc.cfg.raw_to_float_depth, sc.cfg.baseline_fx, sc.cfg.cell, sc.depth_a, sc.cfactor, **kw)
    od.build((sc.depth[0], sc.normals[0], sc.color[0]), frame, num_scales=num_scales, **build_kw)
    return od


def test_pyramid_structure(pair):
    sc, _, frame = pair
    od = make(sc, frame)
    h, w = sc.cfg.height, sc.cfg.width
    for s, L in enumerate(od.levels):
        for img in (L["base"], L["tracked"]):
            d, n, c = img
            assert d.shape == n.shape == c.shape == (h >> s, w >> s) and d.dtype == np.float32 and c.dtype == np.uint8
        if s == 0:
            continue
        # every downsampled depth is one of the four depths of its 2x2 block (the one closest to their mean), with its normal
        pd, pn, pc = od.levels[s - 1]["base"]
        d, n, c = L["base"]
        hh, ww = d.shape
        blocks = np.stack([pd[dy:2 * hh:2, dx:2 * ww:2] for dy in (0, 1) for dx in (0, 1)])
        nblocks = np.stack([pn[dy:2 * hh:2, dx:2 * ww:2] for dy in (0, 1) for dx in (0, 1)])
        valid = d > 0
        assert np.array_equal(valid, (blocks > 0).any(0))
        hit = (blocks == d[None]) & (nblocks == n[None])
        assert hit.any(0)[valid].all()
        mean = np.where(blocks > 0, blocks, 0).sum(0) / np.maximum((blocks > 0).sum(0), 1)
        best = np.where(blocks > 0, np.abs(blocks - mean[None]), np.inf).min(0)
        assert np.all(np.abs(d - mean)[valid] <= best[valid] + 1e-6)
        # intensity: the rounded mean of the block
        cm = np.stack([pc[dy:2 * hh:2, dx:2 * ww:2] for dy in (0, 1) for dx in (0, 1)]).astype(np.float64).mean(0)
        assert np.abs(c.astype(np.float64) - cm).max() <= 0.5 + 1e-3
    # level 0 of the tracked frame: calibrated depth of every valid raw pixel, intensity = luma (or one below: the reference
    # truncates 255 * (v / 255) twice, cuda_image_processing.cu:203 and cuda_buffer.cu:89)
    d0, _, c0 = od.levels[0]["tracked"]
    assert np.array_equal(d0 > 0, (frame[0] & 0x8000) == 0)
    diff = frame[2][..., 3].astype(np.int32) - c0.astype(np.int32)
    assert diff.min() >= 0 and diff.max() <= 2


def test_without_pyramid_level_0(pair):
    sc, _, frame = pair
    od = make(sc, frame, use_pyramid_level_0=False)
    assert od.first_scale == 1 and "tracked" not in od.levels[0]
    full = make(sc, frame)
    d1, n1, c1 = od.levels[1]["tracked"]
    d2, n2, c2 = full.levels[1]["tracked"]
    assert d1.shape == d2.shape and np.array_equal(d1 > 0, d2 > 0)
    # (CalibrateAndDownsample picks among raw-calibrated depths, Downsample among already calibrated ones: same values up to
    #  the cfactor cell quirk; the colour differs by the intermediate rounding only)
    assert np.allclose(d1, d2, rtol=1e-5) and np.abs(c1.astype(int) - c2.astype(int)).max() <= 1


def test_pose_jacobians_numerically(pair):
    """d residual / d pose (global_T_frame * exp(delta) convention, kernel_opt_pose.cu:45-94, 144-190) against central
    differences; the analytic form keeps the association (and with it the measured depth / the sampled texels) fixed, so the
    comparison uses pixels whose projection stays inside one tracked pixel for the perturbation."""
    sc, true_rel, frame = pair
    od = make(sc, frame)
    p0 = S.se3_mul(true_rel, S.se3_exp([0.003, -0.002, 0.001, 0.001, 0.002, -0.001]))
    e0 = od._eval(1, p0, True)
    eps = 2e-4
    plus = [od._eval(1, S.se3_mul(p0, S.se3_exp(np.eye(6)[i] * eps)), False) for i in range(6)]
    minus = [od._eval(1, S.se3_mul(p0, S.se3_exp(-np.eye(6)[i] * eps)), False) for i in range(6)]
    v = e0["visible"].copy()
    for e in plus + minus:
        v &= e["visible"]
    with np.errstate(invalid="ignore"):   # (pixels outside `v` carry NaN / inf)
        num = np.stack([(plus[i]["raw_depth"] - minus[i]["raw_depth"]) / (2 * eps) for i in range(6)])
        ana = e0["Jd"]
        err = np.abs(num - ana)[:, v].max(0) / (np.abs(ana)[:, v].max(0) + 1e-9)
    assert np.median(err) < 0.05 and v.sum() > 1000, (np.median(err), v.sum())


def test_tracking_recovers_the_rendered_motion(pair):
    sc, true_rel, frame = pair
    od = make(sc, frame)
    est, iterations, chose = od.track(IDENT, S.se3_exp([0.01, 0, 0, 0, 0, 0]))
    e0, e1 = S.pose_error(IDENT, true_rel), S.pose_error(est, true_rel)
    assert e1[0] < 0.5 * e0[0] and e1[1] < 0.5 * e0[1], (e0, e1)
    assert all(1 <= it <= 30 for it in iterations) and chose[2] in (0, 1) and chose[0] in (0, 1)
    # the cost at the result is lower than at the start on the finest level
    assert od.cost(0, est)[1] / od.cost(0, est)[0] < od.cost(0, IDENT)[1] / max(od.cost(0, IDENT)[0], 1)


def assert_same_up_to_ties(prev_depth, depth_a, depth_b, what):
    """Two downsampled depth images of the same finer level: wherever they differ, both picks are (to rounding) equally close
    to the block mean.  On planar surfaces the four depths of a 2x2 block are pairwise symmetric about their mean, so the
    reference's "closest to the average" (kernel_downsample.cu:72-90) is decided by the last bits of an approximate division
    (-use_fast_math) -- which a CPU restatement cannot reproduce and does not need to."""
    hh, ww = depth_a.shape
    blocks = np.stack([prev_depth[dy:2 * hh:2, dx:2 * ww:2] for dy in (0, 1) for dx in (0, 1)]).astype(np.float64)
    valid = blocks > 0
    mean = np.where(valid, blocks, 0).sum(0) / np.maximum(valid.sum(0), 1)
    assert np.array_equal(depth_a > 0, depth_b > 0), what
    ok = depth_b > 0
    da, db = np.abs(depth_a - mean)[ok], np.abs(depth_b - mean)[ok]
    assert np.all(np.abs(da - db) <= 4e-6 * mean[ok]), (what, np.max(np.abs(da - db) / mean[ok]))


@pytest.mark.parametrize("tag,gm", [("", False), ("_gradmag", True)])
def test_oracle_matches_reference_golden(pair, tag, gm):
    assert os.path.exists(GOLDEN), "tests/golden/tiny_odometry.npz is committed (tools/make_golden.py --odometry-only on a B200)"
    g = np.load(GOLDEN)
    sc, true_rel, frame = pair
    assert int(frame[0].astype(np.uint64).sum()) + int(frame[2].astype(np.uint64).sum()) == int(g["frame_checksum"])
    S_ = int(g["num_scales"])
    od = make(sc, frame, num_scales=S_, use_gradmag=gm)
    for s in range(S_):
        for wn in ("base", "tracked"):
            gd, gn, gc = g[f"{wn}{s}_depth{tag}"], g[f"{wn}{s}_normals{tag}"], g[f"{wn}{s}_color{tag}"]
            if s == 0:
                d, n, c = od.levels[0][wn]
                prev = None
            else:
                # one level of the oracle's pyramid builder on the REFERENCE's finer level
                prev = (g[f"{wn}{s - 1}_depth{tag}"], g[f"{wn}{s - 1}_normals{tag}"], g[f"{wn}{s - 1}_color{tag}"])
                d, n, c = O.downsample(*prev)
            if gm and s == 0:
                # Sobel magnitude: sqrt + a float product truncated to u8 under -use_fast_math; one level off on a few pixels
                dc = np.abs(c.astype(int) - gc.astype(int))
                assert dc.max() <= 1 and np.mean(dc != 0) < 0.02, (wn, s, dc.max(), np.mean(dc != 0))
            else:
                assert np.array_equal(c, gc), (wn, s, np.mean(c != gc))
            valid = gd > 0
            assert np.array_equal(d > 0, valid)
            if s == 0:
                assert np.allclose(d[valid], gd[valid], rtol=2e-6, atol=0)
                assert np.array_equal(np.where(valid, n, 0), gn), (wn, s)
            else:
                assert_same_up_to_ties(prev[0], d, gd, (wn, s))
                same = valid & (d == gd)
                assert same.mean() > 0.5 and np.array_equal(n[same], gn[same]), (wn, s)
            # the evaluations below run on the reference's own images of this level
            od.levels[s][wn] = (gd, gn, gc)
        H, b, cnt, total = od.coeffs(s, g["true_rel"])
        assert abs(cnt - int(g[f"count{s}{tag}"])) <= max(2, 3e-4 * cnt), (s, cnt, int(g[f"count{s}{tag}"]))
        tol = 5e-3 if gm else 2e-3
        assert rel(H, g[f"H{s}{tag}"]) < tol and rel(b, g[f"b{s}{tag}"]) < tol, (s, rel(H, g[f"H{s}{tag}"]), rel(b, g[f"b{s}{tag}"]))
        assert abs(total - float(g[f"sum{s}{tag}"])) < tol * float(g[f"sum{s}{tag}"])
        for i, pose in enumerate((g["true_rel"], g["off"])):
            c_, cost = od.cost(s, pose)
            assert abs(c_ - int(g[f"cost_counts{s}{tag}"][i])) <= max(3, 3e-4 * c_)
            assert abs(cost - float(g[f"cost_costs{s}{tag}"][i])) < tol * float(g[f"cost_costs{s}{tag}"][i])
    if not gm:
        # the whole optimisation on the reference's pyramids: same branch decisions, result within the drift this iteration
        # shows between two runs of the reference itself (+ 1 mm: 30 capped, non-settling iterations per level amplify
        # last-bit differences)
        est, iterations, chose = od.track(IDENT, g["init2"])
        assert chose == list(g["chose_initial"])
        noise = max(S.pose_error(g["est"], g["est_rerun"]))
        dt, dr = S.pose_error(est, g["est"])
        assert dt < 1e-3 + 3 * noise and dr < 1e-3 + 3 * noise, (dt, dr, noise)


def test_edge_cases_empty_and_identical_frames(pair):
    """No usable pixel at all -> no residuals, zero normal equations, the estimate stays where it was (the reference's LDLT of a
    zero matrix yields a zero update, which counts as converged: one iteration per level).  Tracking a keyframe against ITSELF
    starts and stays at the identity."""
    sc, true_rel, frame = pair
    depth, normals, color = frame
    empty = (np.full_like(depth, 65535), np.zeros_like(normals), color)
    od = make(sc, empty)
    for s in range(od.num_scales):
        H, b, cnt, total = od.coeffs(s, IDENT)
        assert cnt == 0 and total == 0.0 and not np.any(H) and not np.any(b)
        assert od.cost(s, IDENT) == (0, 0.0)
    start = S.se3_exp([0.01, 0.0, -0.01, 0.0, 0.002, 0.0])
    est, iterations, chose = od.track(start, start)
    assert np.allclose(est, start, atol=1e-7) and iterations == [1, 1, 1]
    same = make(sc, (sc.depth[0], sc.normals[0], sc.color[0]))
    est, iterations, _ = same.track(IDENT, IDENT)
    dt, dr = S.pose_error(est, IDENT)
    assert dt < 2e-4 and dr < 2e-4, (dt, dr)
    # at the identity every valid interior pixel of the keyframe is associated with itself
    H, b, cnt, total = same.coeffs(0, IDENT)
    d0 = same.levels[0]["base"][0]
    assert cnt >= 2 * 0.9 * (d0[:-1, :-1] > 0).sum()
