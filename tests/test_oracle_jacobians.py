"""Known-answer tests of the oracle's residuals / Jacobians against the reference's SYMBOLIC residual
definitions (applications/badslam/scripts/jacobians_derivation.py:170-302), evaluated numerically in fp64:

  depth residual        inv_sigma * dot(n, global_T_frame * exp(hat(T)) * unproject(x, y, d) - s)          (:219-232)
  ... wrt surfel pos    inv_sigma * dot(n, global_point - (s + t n))                                         (:238-246)
  descriptor residual   interp_bilinear(I, Project(SE3Inverse(exp(hat(T))) * local_surfel_pos)) - d          (:286-301)

The symbolic script differentiates exactly these expressions at T = 0 / t = 0; here the same derivatives are
taken by central differences, and compared with the analytic Jacobians the oracle restates from
kernel_opt_pose.cu:45-142 and kernel_opt_geometry.cu:118-231.
"""
import numpy as np
import pytest

from badslam_b200 import scene as S
from oracle import cpu_oracle as O


def _pairs(scene, orc, k, count=40, seed=0):
    rng = np.random.default_rng(seed)
    idx = rng.permutation(scene.num_surfels)
    out = []
    for i in idx:
        s8 = scene.surfels[:8, i].copy()
        flags, r, Jp, Jg = orc.pair_residuals(k, s8, scene.poses_true[k])
        if flags == 3:
            out.append((i, s8, r.copy(), Jp.copy(), Jg.copy(), orc.last_debug.copy()))
        if len(out) >= count:
            break
    assert len(out) >= count // 2
    return out


def _se3_exp64(a):
    a = np.asarray(a, np.float64)
    ups, om = a[:3], a[3:]
    th = np.linalg.norm(om)
    Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-12:
        R, V = np.eye(3) + Om, np.eye(3) + 0.5 * Om
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = R, V @ ups
    return M


@pytest.mark.parametrize("k", [0, 2])
def test_depth_residual_pose_and_position_jacobians(small_scene, k):
    sc = small_scene
    orc = O.Oracle(sc)
    G = S.se3_matrix(sc.poses_true[k])                      # global_T_frame
    fx, fy, cx, cy = [float(v) for v in sc.depth_K]
    for (i, s8, r, Jp, Jg, dbg) in _pairs(sc, orc, k):
        px, py, d, inv_sigma = [float(v) for v in dbg[:4]]
        local_point = np.array([d * (px + 0.5 - cx) / fx, d * (py + 0.5 - cy) / fy, d, 1.0])   # unproject(x, y, d)
        n = S.unpack_surfel_normal(s8[3:4].view(np.uint32))[0].astype(np.float64)
        s = s8[:3].astype(np.float64)

        def res(delta):
            return inv_sigma * n @ ((G @ _se3_exp64(delta) @ local_point)[:3] - s)

        assert abs(res(np.zeros(6)) - r[0]) < 2e-3 * max(1.0, abs(r[0]))
        eps = 1e-6
        num = np.array([(res(eps * np.eye(6)[j]) - res(-eps * np.eye(6)[j])) / (2 * eps) for j in range(6)])
        assert np.allclose(num, Jp[0], rtol=2e-3, atol=2e-3 * np.max(np.abs(num))), (i, num, Jp[0])

        # wrt displacement t along the normal (jacobians_derivation.py:238-246): d/dt = -inv_sigma
        gp = (G @ local_point)[:3]
        num_t = (inv_sigma * n @ (gp - (s + eps * n)) - inv_sigma * n @ (gp - (s - eps * n))) / (2 * eps)
        assert abs(num_t - Jg[0, 0]) < 1e-3 * abs(num_t)


def _bilinear_cell(luma, x, y):
    """interp_bilinear on the four texels of the cell (x, y) falls in (pixel-corner coords), as an analytic function."""
    ix = int(max(0.0, x - 0.5))
    iy = int(max(0.0, y - 0.5))
    h, w = luma.shape
    t = lambda i, j: luma[min(max(j, 0), h - 1), min(max(i, 0), w - 1)] / 255.0
    tl, tr, bl, br = t(ix, iy), t(ix + 1, iy), t(ix, iy + 1), t(ix + 1, iy + 1)

    def f(xx, yy):
        a, b = xx - 0.5 - ix, yy - 0.5 - iy
        return (1 - a) * (1 - b) * tl + a * (1 - b) * tr + (1 - a) * b * bl + a * b * br
    return f


@pytest.mark.parametrize("k", [1, 3])
def test_descriptor_residual_pose_and_position_jacobians(small_scene, k):
    sc = small_scene
    orc = O.Oracle(sc)
    G = S.se3_matrix(sc.poses_true[k])
    F = np.linalg.inv(G)                                    # frame_T_global
    fx, fy, cx, cy = [float(v) for v in sc.color_K]
    luma = sc.color[k][..., 3]
    checked = 0
    for (i, s8, r, Jp, Jg, dbg) in _pairs(sc, orc, k, count=60, seed=1):
        ccx, ccy, t1x, t1y, t2x, t2y = [float(v) for v in dbg[4:10]]
        # stay away from cell borders, where the piecewise-bilinear function is not differentiable
        def interior(x, y):
            return min((x - 0.5) % 1.0, 1 - (x - 0.5) % 1.0, (y - 0.5) % 1.0, 1 - (y - 0.5) % 1.0) > 0.02
        if not (interior(ccx, ccy) and interior(t1x, t1y) and interior(t2x, t2y)):
            continue
        checked += 1
        Bc, B1, B2 = _bilinear_cell(luma, ccx, ccy), _bilinear_cell(luma, t1x, t1y), _bilinear_cell(luma, t2x, t2y)
        s = np.append(s8[:3].astype(np.float64), 1.0)
        n = S.unpack_surfel_normal(s8[3:4].view(np.uint32))[0].astype(np.float64)
        ls0 = (F @ s)[:3]
        p0 = np.array([fx * ls0[0] / ls0[2] + cx, fy * ls0[1] / ls0[2] + cy])

        def proj_pose(delta):     # Project(SE3Inverse(exp(hat(T))) * local_surfel_pos)
            ls = (np.linalg.inv(_se3_exp64(delta)) @ np.append(ls0, 1.0))[:3]
            return np.array([fx * ls[0] / ls[2] + cx, fy * ls[1] / ls[2] + cy])

        def proj_t(t):            # Project(frame_T_global * (surfel_pos + t n))
            ls = (F @ np.append(s[:3] + t * n, 1.0))[:3]
            return np.array([fx * ls[0] / ls[2] + cx, fy * ls[1] / ls[2] + cy])

        # the reference's approximation: all three sample points move like the centre point (cost_function.cuh:245-248)
        def desc(p, which):
            dp = p - p0
            Bt, tx, ty = (B1, t1x, t1y) if which == 1 else (B2, t2x, t2y)
            return 180.0 * (Bt(tx + dp[0], ty + dp[1]) - Bc(ccx + dp[0], ccy + dp[1]))

        eps = 1e-6
        for which, row in ((1, 1), (2, 2)):
            num = np.array([(desc(proj_pose(eps * np.eye(6)[j]), which) - desc(proj_pose(-eps * np.eye(6)[j]), which)) / (2 * eps)
                            for j in range(6)])
            scale = max(np.max(np.abs(num)), 1e-3)
            assert np.allclose(num, Jp[row], rtol=5e-3, atol=5e-3 * scale), (i, which, num, Jp[row])
            num_t = (desc(proj_t(eps), which) - desc(proj_t(-eps), which)) / (2 * eps)
            assert abs(num_t - Jg[row, 0]) < 5e-3 * max(abs(num_t), 1e-2 * scale), (i, which, num_t, Jg[row, 0])
            assert Jg[row, row] == -1.0        # d residual / d descriptor
        # raw residual value: 180 (I(t_i) - I(c)) - d_i with fp32 / 8-bit filter weights
        assert abs(r[1] - (180.0 * (B1(t1x, t1y) - Bc(ccx, ccy)) - float(s8[6]))) < 0.5
    assert checked >= 10
