"""CPU-only: keyframe preprocessing (bad_slam.cc:692-765, cuda_depth_processing.cu, cuda_image_processing.cu).

* the oracle's stages (oracle/preprocess_oracle.c) against closed forms and against the independent numpy restatement of the
  Keyframe constructor's preprocessing in badslam_b200/scene.py (normals -> radii);
* the tile program the CUDA kernel runs (badslam_b200/csrc/preprocess_tile.cuh), executed on the host by
  tests/harness/preprocess_host.cpp, against the oracle's whole-image passes: same pixels dropped, same values up to the
  rounding of one float operation (the harness is compiled without FMA contraction, the luma of the oracle follows nvcc's).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from badslam_b200 import scene as S
from oracle import cpu_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
UNKNOWN = 65535


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "harness", "preprocess_host.cpp")
    hdr = os.path.join(HERE, "..", "badslam_b200", "csrc", "preprocess_tile.cuh")
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libpreprocess_host.so")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", so])
    lib = C.CDLL(so)
    lib.harness_float_to_half.restype = C.c_uint16
    lib.harness_float_to_half.argtypes = [C.c_float]
    return lib


@pytest.fixture(scope="module")
def small():
    sc = S.make_scene(S.config_by_name("small"))
    return sc, O.Oracle(sc)


def run_harness(lib, sc, orc, raw, rgb, sigma_xy=1.5, sigma_inv=0.005, radius_factor=2.0, max_depth=3.0):
    m = orc.model
    h, w = raw.shape
    depth, normals, radius = (np.zeros_like(raw) for _ in range(3))
    rgba = np.zeros((h, w, 4), np.uint8)
    mm = np.zeros(2, np.float32)
    K = (C.c_float * 4)(*[float(v) for v in sc.depth_K])
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    tiles = lib.harness_preprocess_frame(
        C.c_int(w), C.c_int(h), K, C.c_float(m.raw_to_float_depth), C.c_float(m.a), C.c_int(m.cell), C.c_int(m.cf_w),
        p(orc.cfactor, C.c_float), C.c_float(sigma_xy), C.c_float(sigma_inv), C.c_float(radius_factor), C.c_float(max_depth),
        p(raw, C.c_uint16), p(depth, C.c_uint16), p(normals, C.c_uint16), p(radius, C.c_uint16),
        C.c_int(w), C.c_int(h), p(np.ascontiguousarray(rgb), C.c_uint8), p(rgba, C.c_uint8), p(mm, C.c_float))
    assert tiles == -(-w // 32) * -(-h // 32)
    return depth, normals, radius, rgba, float(mm[0]), float(mm[1])


def s8_pair(n16):
    return (n16 & 0xff).astype(np.int8).astype(np.int32), (n16 >> 8).astype(np.int8).astype(np.int32)


def test_float_to_half_matches_ieee(harness):
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        np.float32(10.0) ** rng.uniform(-9, 5.2, 20000).astype(np.float32),
        np.array([0, 1, 65504, 65519.99, 65520, 1e9, 2.0 ** -14, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 5.96e-8, 6.1e-5,
                  1.0009765625, 1.00048828125, 1.00146484375, np.inf], np.float32)])
    lib = O.lib()
    lib.orc_float_to_half.restype = C.c_uint16
    lib.orc_float_to_half.argtypes = [C.c_float]
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    got_o = np.array([lib.orc_float_to_half(float(v)) for v in vals], np.uint16)
    got_h = np.array([harness.harness_float_to_half(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got_o, want) and np.array_equal(got_h, want)


def test_bilateral_filter_properties(small):
    sc, orc = small
    h, w = sc.cfg.height, sc.cfg.width
    # a fronto-parallel plane is a fixed point (up to the truncation of the float result): |out - in| <= 1 raw unit
    flat = np.full((h, w), 2000, np.uint16)
    out = orc.bilateral_filter(flat)
    assert np.all(np.abs(out.astype(np.int32) - 2000) <= 1)
    # depth cut-off and missing measurements (cuda_depth_processing.cu:54-58): unknown, and not used as samples
    raw = flat.copy()
    raw[10, 10] = 0
    raw[20, 20] = 20000
    out = orc.bilateral_filter(raw, max_depth_raw=15000)
    assert out[10, 10] == UNKNOWN and out[20, 20] == UNKNOWN
    assert abs(int(out[10, 11]) - 2000) <= 1            # the hole is skipped, not averaged in
    # an out-of-range neighbour IS a sample (only the centre is cut off), but its weight is negligible: 1/2 m vs 1/20 m
    assert abs(int(out[20, 21]) - 2000) <= 1
    # edge preservation: a 10 cm step stays a step (sigma_inv_depth = 0.005 1/m; the step is 0.024 1/m)
    step = flat.copy()
    step[:, w // 2:] = 2100
    out = orc.bilateral_filter(step)
    assert abs(int(out[h // 2, w // 2 - 1]) - 2000) <= 1 and abs(int(out[h // 2, w // 2]) - 2100) <= 1
    # noise is reduced on a plane
    rng = np.random.default_rng(1)
    noisy = (2000 + rng.normal(0, 3, (h, w))).round().astype(np.uint16)
    out = orc.bilateral_filter(noisy)
    assert out[8:-8, 8:-8].astype(np.float64).std() < 0.45 * noisy[8:-8, 8:-8].astype(np.float64).std()
    # radius 0 (radius_factor * sigma_xy + 0.5 < 1): identity up to the truncation
    out = orc.bilateral_filter(noisy, sigma_xy=0.2, radius_factor=2.0)
    assert np.all(np.abs(out.astype(np.int32) - noisy.astype(np.int32)) <= 1)


def test_normals_and_radii_match_the_numpy_restatement(small):
    """scene.preprocess_depth is an independent (vectorised numpy) restatement of keyframe.cc:96-144."""
    sc, orc = small
    raw, _ = S.raw_frame(sc, 0)
    a = orc.bilateral_filter(raw)
    d1, n1 = orc.compute_normals(a)
    rad, d2 = orc.compute_radii(d1)
    d2_np, n_np, rad_np, _, _ = S.preprocess_depth(sc.cfg, sc.depth_K, a, orc.cfactor, sc.depth_a)
    assert np.array_equal(d2, d2_np)
    valid1 = (d1 & 0x8000) == 0
    ax, ay = s8_pair(n1[valid1])
    bx, by = s8_pair(n_np[valid1])
    assert max(np.abs(ax - bx).max(), np.abs(ay - by).max()) <= 1 and np.mean((ax != bx) | (ay != by)) < 2e-3
    valid2 = (d2 & 0x8000) == 0
    ra, rb = rad[valid2].view(np.float16).astype(np.float64), rad_np[valid2].view(np.float16).astype(np.float64)
    assert np.all(np.abs(ra - rb) <= 2.0 ** -9 * rb) and np.mean(ra != rb) < 2e-3       # <= 2 half ulps
    # structure: a one-pixel border and every pixel with an invalid 4-neighbour is dropped by the normals stage
    assert np.all(d1[0] == UNKNOWN) and np.all(d1[:, 0] == UNKNOWN) and np.all(d1[-1] == UNKNOWN) and np.all(d1[:, -1] == UNKNOWN)
    inv = (a & 0x8000) != 0
    nb_inv = inv.copy()
    nb_inv[1:] |= inv[:-1]; nb_inv[:-1] |= inv[1:]; nb_inv[:, 1:] |= inv[:, :-1]; nb_inv[:, :-1] |= inv[:, 1:]
    assert np.array_equal(valid1[1:-1, 1:-1], ~nb_inv[1:-1, 1:-1])
    assert np.array_equal(d1[valid1], a[valid1])
    # the plane normals are recovered from a noise-free frame (the scene's own normals come from the unfiltered rendering; the
    # filter's truncation to whole raw units tilts the finite differences by a few s8 steps)
    clean, _ = S.raw_frame(sc, 0, noise_raw=0.0, hole_fraction=0.0, far_fraction=0.0)
    dc, nc = orc.compute_normals(orc.bilateral_filter(clean))
    both = ((dc & 0x8000) == 0) & ((sc.depth[0] & 0x8000) == 0)
    ax, ay = s8_pair(nc[both])
    bx, by = s8_pair(sc.normals[0][both])
    assert both.mean() > 0.5 and np.median(np.hypot(ax - bx, ay - by)) < 12


def test_radius_is_the_squared_distance_to_the_nearest_neighbour(small):
    sc, orc = small
    h, w = sc.cfg.height, sc.cfg.width
    flat = np.full((h, w), 2000, np.uint16)
    d1, _ = orc.compute_normals(flat)
    rad, d2 = orc.compute_radii(d1)
    valid = (d2 & 0x8000) == 0
    assert valid[2:-2, 2:-2].all() and not valid[1].any()       # 2-pixel frame: border, then pixels next to the border
    want = (2.0 / float(sc.depth_K[0])) ** 2                     # neighbouring rays are depth / fx apart at 2 m
    got = rad[valid].view(np.float16).astype(np.float64)
    assert np.all(np.abs(got - want) <= 2.0 ** -10 * want)
    mn, mx = C.c_float(), C.c_float()
    orc.lib.orc_compute_min_max_depth(C.c_int(w), C.c_int(h), C.c_float(orc.model.raw_to_float_depth),
                                      d2.ctypes.data_as(C.POINTER(C.c_uint16)), C.byref(mn), C.byref(mx))
    assert abs(mn.value - 2.0) < 1e-6 and abs(mx.value - 2.0) < 1e-6
    empty = np.full((h, w), UNKNOWN, np.uint16)
    orc.lib.orc_compute_min_max_depth(C.c_int(w), C.c_int(h), C.c_float(1e-3), empty.ctypes.data_as(C.POINTER(C.c_uint16)),
                                      C.byref(mn), C.byref(mx))
    assert mn.value == float("inf") and mx.value == 0.0


def test_brightness_follows_the_compiled_contraction():
    lib = O.lib()
    rng = np.random.default_rng(2)
    rgb = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)
    rgb[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 255], [128, 128, 128]]
    rgba = np.zeros((64, 64, 4), np.uint8)
    lib.orc_compute_brightness(C.c_int(64), C.c_int(64), rgb.ctypes.data_as(C.POINTER(C.c_uint8)), rgba.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(rgba[..., :3], rgb)
    r, g, b = [rgb[..., i].astype(np.float64) for i in range(3)]
    exact = 0.299 * r + 0.587 * g + 0.114 * b + 0.5
    assert np.all(np.abs(rgba[..., 3].astype(np.float64) - np.floor(exact)) <= 1)
    assert np.mean(rgba[..., 3] != np.floor(exact)) < 1e-2
    assert list(rgba[0, :2, 3]) == [0, 255]


@pytest.mark.parametrize("radius_factor,sigma_xy", [(2.0, 1.5), (1.0, 1.0), (2.0, 0.2)])
def test_tile_program_matches_the_whole_image_oracle(harness, small, radius_factor, sigma_xy):
    """The fused tile program (halo staging, three stages in shared memory) against five whole-image passes."""
    sc, orc = small
    raw, rgb = S.raw_frame(sc, 1)
    raw[:40, :40] = 0                      # an empty tile corner
    raw[100:103, :] = 0                    # a gap crossing tile borders
    want = orc.preprocess_frame(raw, rgb, sigma_xy=sigma_xy, radius_factor=radius_factor)
    got = run_harness(harness, sc, orc, raw, rgb, sigma_xy=sigma_xy, radius_factor=radius_factor)
    wd, wn, wr, wc, wmin, wmax = want
    gd, gn, gr, gc, gmin, gmax = got
    # identical validity; both are IEEE float programs of the same expressions, so the values agree bit for bit
    assert np.array_equal(gd, wd)
    valid = (wd & 0x8000) == 0
    assert valid.mean() > 0.5
    assert np.array_equal(gn, wn)
    assert np.array_equal(gr[valid], wr[valid]) and np.all(gr[~valid] == 0)
    assert gmin == wmin and gmax == wmax and 0 < gmin < gmax < 3.0
    assert np.array_equal(gc[..., :3], wc[..., :3])
    dl = np.abs(gc[..., 3].astype(np.int32) - wc[..., 3].astype(np.int32))
    assert dl.max() <= 1 and np.mean(dl != 0) < 1e-2       # FMA contraction of the luma (oracle) vs none (host build)


def test_tile_program_on_ragged_and_tiny_images(harness):
    """Image sizes that are not multiples of the tile, smaller than a tile, and smaller than the filter window."""
    for (w, h) in [(70, 45), (33, 31), (8, 5), (3, 3), (1, 1)]:
        sc = S.blank_scene(w, h)
        raw, rgb = S.random_raw_frame(w, h, seed=w * 100 + h)
        orc = O.Oracle(sc)
        want = orc.preprocess_frame(raw, rgb)
        got = run_harness(harness, sc, orc, raw, rgb)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (w, h)
        valid = (want[0] & 0x8000) == 0
        assert np.array_equal(got[2][valid], want[2][valid]), (w, h)
        assert got[4] == want[4] and got[5] == want[5], (w, h)
        if min(w, h) < 5:
            assert not valid.any()      # nothing survives a 2-pixel frame


def test_tile_program_is_clean_under_address_sanitizer(tmp_path):
    """Every global / shared-memory index of the tile program, on ragged sizes and filter radii 0 / 3 / 16, with exactly sized
    buffers under ASan + UBSan."""
    exe = str(tmp_path / "asan_test")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-x", "c++",
           os.path.join(HERE, "harness", "preprocess_host.cpp"), os.path.join(HERE, "harness", "preprocess_asan_main.cpp"), "-o", exe]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0:
        pytest.skip("no sanitizer runtime for this g++: " + built.stderr[-200:])
    run = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert run.returncode == 0 and "runtime error" not in run.stderr and "AddressSanitizer" not in run.stderr, run.stderr[-2000:]
    assert "320x240 sigma 8.0: 80 tiles" in run.stdout


def test_preprocessed_raw_frames_feed_surfel_creation_and_bundle_adjustment(small):
    """The chain BadSlam::ProcessFrame -> CreateKeyframe -> RunBundleAdjustment takes (bad_slam.cc:640-1010), in the oracle: noisy raw
    frames -> preprocessing -> keyframes -> CreateSurfelsForKeyframe -> 3 BA iterations improve the relative poses."""
    import copy
    sc, orc0 = small
    K = sc.cfg.num_keyframes
    outs = [orc0.preprocess_frame(*S.raw_frame(sc, k, noise_raw=1.0), max_depth=6.0) for k in range(K)]
    sc2 = copy.copy(sc)
    sc2.depth, sc2.normals, sc2.radius, sc2.color = (np.stack([o[i] for o in outs]) for i in range(4))
    sc2.min_depth = np.array([o[4] for o in outs], np.float32)
    sc2.max_depth = np.array([o[5] for o in outs], np.float32)
    assert np.all(sc2.min_depth > 0) and np.all(sc2.max_depth <= 6.0)
    sc2.surfels = np.zeros((17, 1 << 17), np.float32)
    sc2.num_surfels = 0
    orc = O.Oracle(sc2)
    created = sum(orc.create_surfels_for_keyframe(k, True) for k in range(K))
    assert created > 20000 and orc.n == created
    r = orc.bundle_adjust(True, True, 3, 3)
    assert r.iterations_done == 3 and r.n_assoc > created

    def rel(P, k):
        return S.se3_mul(S.se3_inverse(P[0]), P[k])
    e_init = max(S.pose_error(rel(sc.poses_init, k), rel(sc.poses_true, k))[0] for k in range(1, K))
    e_ba = max(S.pose_error(rel(orc.poses, k), rel(sc.poses_true, k))[0] for k in range(1, K))
    assert e_ba < 0.8 * e_init


def test_preprocessing_matches_reference_cuda_golden():
    """tests/golden/tiny_preprocess.npz: outputs of the reference's own preprocessing kernels (tools/make_golden.py::
    golden_preprocess, produced on the GPU box).  The reference is a -use_fast_math build: the filtered depth may differ from the
    IEEE oracle by one raw unit on a small fraction of the pixels; validity, luma and everything computed from agreeing depths
    must match."""
    path = os.path.join(HERE, "golden", "tiny_preprocess.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture not generated yet (needs the reference kernels on a GPU: tools/make_golden.py --preprocess-only)")
    g = np.load(path)
    sc = S.make_scene(S.config_by_name("tiny"))
    sc.depth_a = 0.02
    sc.cfactor = (2e-3 * np.random.default_rng(5).random(sc.cfactor.shape)).astype(np.float32)
    raw, rgb = S.raw_frame(sc, int(g["kf"]))
    assert int(raw.astype(np.uint64).sum()) == int(g["raw_checksum"])
    d, n, r, c, mn, mx = O.Oracle(sc).preprocess_frame(raw, rgb)
    valid = (g["depth"] & 0x8000) == 0
    assert np.array_equal((d & 0x8000) == 0, valid)
    dd = np.abs(d[valid].astype(np.int32) - g["depth"][valid].astype(np.int32))
    assert dd.max() <= 1 and np.mean(dd != 0) < 3e-2
    assert np.array_equal(c[..., 3], g["luma"])
    same = valid & (d == g["depth"])
    nb = same.copy()
    nb[1:] &= same[:-1]; nb[:-1] &= same[1:]; nb[:, 1:] &= same[:, :-1]; nb[:, :-1] &= same[:, 1:]
    ax, ay = s8_pair(n[nb])
    bx, by = s8_pair(g["normals"][nb])
    assert max(np.abs(ax - bx).max(), np.abs(ay - by).max()) <= 1 and np.mean((ax != bx) | (ay != by)) < 2e-2
    ra, rb = r[nb].view(np.float16).astype(np.float64), g["radius"][nb].view(np.float16).astype(np.float64)
    assert np.all(np.abs(ra - rb) <= np.maximum(2.0 ** -9 * rb, 2.0 ** -23))   # two half ulps, normal or subnormal
    assert abs(mn - float(g["min_depth"])) <= 1.5e-3 and abs(mx - float(g["max_depth"])) <= 1.5e-3
