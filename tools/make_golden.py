"""Generates tests/golden/*.npz from the REFERENCE's own CUDA kernels (oracle/_ref) on the GPU box.

    gpurun -- python tools/make_golden.py        # writes gpurun_out/golden/<scene>.npz
    cp gpurun_out/golden/*.npz tests/golden/      # commit

The reference holds no golden vectors for this path; these fixtures are outputs of the reference
implementation itself on the seeded synthetic scenes of badslam_b200/scene.py, so that the CPU-only test suite can
pin the oracle (and the GPU suite the CUDA path) against the reference without /root/reference being present.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from badslam_b200.scene import config_by_name, make_scene  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def golden_for(name, use_depth=True, use_desc=True, tag=""):
    sc = make_scene(config_by_name(name))
    K = sc.cfg.num_keyframes
    out = {"scene": name, "use_depth": use_depth, "use_desc": use_desc, "num_surfels": sc.num_surfels,
           "surfel_checksum": float(np.sum(sc.surfels[:3, :sc.num_surfels].astype(np.float64))),
           "depth_checksum": int(sc.depth.astype(np.uint64).sum())}
    ref = ref_cuda.RefDirectBA(sc, use_depth, use_desc)
    H, b, cnt, cost = [], [], [], []
    for k in range(K):
        h_, b_, c_, s_ = ref.pose_coeffs(k, sc.poses_init[k])
        H.append(h_); b.append(b_); cnt.append(c_); cost.append(s_)
    out.update(pose_H=np.array(H), pose_b=np.array(b), pose_count=np.array(cnt), pose_cost=np.array(cost))
    est, its, conv = [], [], []
    for k in range(K):
        p, i, c = ref.estimate_frame_pose(k, sc.poses_init[k])
        est.append(p); its.append(i); conv.append(c)
    out.update(efp_pose=np.array(est), efp_iterations=np.array(its), efp_converged=np.array(conv))
    ref.update_activation()
    out["activation_flags"] = np.packbits(ref.active())
    ref.optimize_geometry_iteration()
    out["geometry_rows"] = ref.surfels()[[0, 1, 2, 3, 6, 7]].copy()
    ref.close()
    ref = ref_cuda.RefDirectBA(sc, use_depth, use_desc)
    r = ref.bundle_adjust(True, True, 3, 3, count_residuals=True, end_tasks=False)
    out.update(ba_poses=ref.poses(), ba_activation=ref.activation(), ba_count=int(r.n_count), ba_cost=float(r.cost),
               ba_pose_iterations=int(r.pose_iterations_total), ba_surfels=ref.surfels()[[0, 1, 2, 6, 7]].copy())
    # run-to-run noise of the reference itself (float atomics): a second identical run
    ref2 = ref_cuda.RefDirectBA(sc, use_depth, use_desc)
    ref2.bundle_adjust(True, True, 3, 3, count_residuals=True, end_tasks=False)
    out["ba_poses_rerun"] = ref2.poses()
    ref.close(); ref2.close()
    os.makedirs("gpurun_out/golden", exist_ok=True)
    np.savez_compressed(f"gpurun_out/golden/{name}{tag}.npz", **out)
    print("wrote", name, tag, "counts", out["pose_count"][:4], "ba_count", out["ba_count"])


def distorted_scene(name="tiny"):
    """Scene with a depth deformation in the raw depth, perturbed camera estimates and a non-zero deformation model."""
    import dataclasses
    sc = make_scene(dataclasses.replace(config_by_name(name), depth_a=0.03, cfactor=0.005))
    sc.depth_K = (np.asarray(sc.depth_K, np.float32) * np.float32([1.003, 0.998, 1.002, 0.997])).astype(np.float32)
    sc.color_K = (np.asarray(sc.color_K, np.float32) * np.float32([0.998, 1.002, 1.001, 0.999])).astype(np.float32)
    a_init = 0.02
    cf_init = (np.random.default_rng(5).standard_normal(sc.cfactor.shape) * 0.003).astype(np.float32)
    return sc, a_init, cf_init


def golden_intrinsics_pcg(name="tiny"):
    """OptimizeIntrinsicsCUDA and the PCG solver's building blocks / short solves from the reference's kernels."""
    sc, a_init, cf_init = distorted_scene(name)
    K, n = sc.cfg.num_keyframes, sc.num_surfels
    out = {"scene": name, "surfel_checksum": float(np.sum(sc.surfels[:3, :n].astype(np.float64)))}
    ref = ref_cuda.RefDirectBA(sc)
    ref.set_depth_params(a_init, cf_init)
    for step in range(2):
        ref.optimize_intrinsics(True, True)
        d, c, a = ref.intrinsics()
        out[f"intr{step}_depth_K"], out[f"intr{step}_color_K"], out[f"intr{step}_a"] = d, c, np.float32(a)
        out[f"intr{step}_cfactor"] = ref.cfactor()
    ref.close()
    for intr in (False, True):
        ref = ref_cuda.RefDirectBA(sc)
        ref.set_depth_params(a_init, cf_init)
        r, M, p, g, scal = ref.pcg_debug(optimize_depth_intrinsics=intr, optimize_color_intrinsics=intr, gauge_keyframe=1)
        tag = "pcgi" if intr else "pcg"
        lo, hi = 6 * (K - 1), 6 * (K - 1) + 3 * n
        for nm, v in (("r", r), ("M", M), ("p", p), ("g", g)):
            out[f"{tag}_{nm}_pose"] = v[:lo].copy()
            out[f"{tag}_{nm}_surfel_sum"] = np.float64(v[lo:hi].astype(np.float64).sum())
            out[f"{tag}_{nm}_surfel_abs"] = np.float64(np.abs(v[lo:hi].astype(np.float64)).sum())
            if intr:
                out[f"{tag}_{nm}_intr"] = v[hi:].copy()
        out[f"{tag}_scalars"] = scal
        res = ref.bundle_adjust_pcg(True, True, intr, intr, 2, 2, 4, 1, end_tasks=False)
        out[f"{tag}_ba_poses"], out[f"{tag}_ba_r_norm"] = ref.poses(), np.float32(res.last_r_norm)
        out[f"{tag}_ba_surfels"] = ref.surfels()[[0, 1, 2, 6, 7]].copy()
        if intr:
            d, c, a = ref.intrinsics()
            out["pcgi_ba_depth_K"], out["pcgi_ba_color_K"], out["pcgi_ba_a"] = d, c, np.float32(a)
        ref2 = ref_cuda.RefDirectBA(sc)
        ref2.set_depth_params(a_init, cf_init)
        ref2.bundle_adjust_pcg(True, True, intr, intr, 2, 2, 4, 1, end_tasks=False)
        out[f"{tag}_ba_poses_rerun"] = ref2.poses()
        ref.close(); ref2.close()
    os.makedirs("gpurun_out/golden", exist_ok=True)
    np.savez_compressed(f"gpurun_out/golden/{name}_intrinsics_pcg.npz", **out)
    print("wrote", name, "intrinsics + pcg", out["intr1_depth_K"], out["pcg_scalars"], out["pcgi_scalars"])


def golden_end_tasks(name="tiny"):
    """PerformBASchemeEndTasks (delete / radius update / compaction) from the reference's kernels on a scene with displaced
    surfels (scene.displace_surfels)."""
    from badslam_b200.scene import displace_surfels
    sc = displace_surfels(make_scene(config_by_name(name)))[0]
    ref = ref_cuda.RefDirectBA(sc)
    deleted = ref.end_tasks()
    rows = ref.surfels()
    out = {"scene": name, "deleted": deleted, "surfels_size": ref.surfels_size(), "rows": rows.copy(),
           "surfel_checksum": float(np.sum(sc.surfels[:3, :sc.num_surfels].astype(np.float64)))}
    ref.close()
    os.makedirs("gpurun_out/golden", exist_ok=True)
    np.savez_compressed(f"gpurun_out/golden/{name}_end_tasks.npz", **out)
    print("wrote", name, "end tasks: deleted", deleted, "of", sc.num_surfels)


def golden_preprocess(name="tiny", kf=1):
    """BadSlam::PreprocessFrame + min / max depth from the reference's kernels (cuda_depth_processing.cu,
    cuda_image_processing.cu) on one noisy raw frame of the scene, with a non-trivial depth deformation."""
    from badslam_b200.scene import raw_frame
    sc = make_scene(config_by_name(name))
    sc.depth_a = 0.02
    sc.cfactor = (2e-3 * np.random.default_rng(5).random(sc.cfactor.shape)).astype(np.float32)
    raw, rgb = raw_frame(sc, kf)
    ref = ref_cuda.RefDirectBA(sc)
    depth, normals, radius, rgba, mn, mx = ref.preprocess_frame(raw, rgb)
    ref.close()
    os.makedirs("gpurun_out/golden", exist_ok=True)
    np.savez_compressed(f"gpurun_out/golden/{name}_preprocess.npz", scene=name, kf=kf, raw_checksum=int(raw.astype(np.uint64).sum()),
                        depth=depth, normals=normals, radius=radius, luma=rgba[..., 3].copy(), min_depth=mn, max_depth=mx)
    print("wrote", name, "preprocess: valid", float(((depth & 0x8000) == 0).mean()), mn, mx)


ODOMETRY_MOTION = [0.02, -0.01, 0.015, 0.01, -0.008, 0.012]   # base_T_frame of the tracked frame (tests/test_oracle_odometry.py)


def golden_odometry(name="tiny", num_scales=3):
    """Image-pair odometry (BadSlam::RunOdometry + TrackFramePairwise) on the reference's own kernels: the pyramids of keyframe 0
    and of a frame rendered 2 cm / 1 degree away from it, the normal equations / costs at two poses on every level, and the
    result of the whole coarse-to-fine optimisation (twice: the reference's own run-to-run difference)."""
    from badslam_b200 import scene as S
    sc = make_scene(config_by_name(name))
    true_rel = S.se3_exp(ODOMETRY_MOTION)
    depth, normals, _, color = S.render_frame(sc, S.se3_mul(sc.poses_true[0], true_rel))
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    init2 = S.se3_exp([0.01, 0.0, 0.0, 0.0, 0.0, 0.0])
    out = {"scene": name, "num_scales": num_scales, "true_rel": true_rel, "init2": init2,
           "frame_checksum": int(depth.astype(np.uint64).sum()) + int(color.astype(np.uint64).sum())}
    for tag, gm in (("", False), ("_gradmag", True)):
        ref = ref_cuda.RefDirectBA(sc)
        est, res = ref.track_frame_pairwise(0, depth, normals, color, ident, init2, num_scales=num_scales, use_gradmag=gm)
        est2, _ = ref.track_frame_pairwise(0, depth, normals, color, ident, init2, num_scales=num_scales, use_gradmag=gm)
        out[f"est{tag}"], out[f"est_rerun{tag}"] = est, est2
        out[f"iterations{tag}"] = np.array(list(res.iterations)[:num_scales])
        out[f"chose_initial{tag}"] = np.array(list(res.chose_initial)[:num_scales])
        off = S.se3_mul(true_rel, S.se3_exp([0.004, -0.003, 0.002, 0.002, 0.001, -0.002]))
        out["off"] = off
        for scale in range(num_scales):
            for which, wn in ((0, "base"), (1, "tracked")):
                d, n, c = ref.odometry_level(which, scale)
                out[f"{wn}{scale}_depth{tag}"], out[f"{wn}{scale}_color{tag}"] = d, c
                out[f"{wn}{scale}_normals{tag}"] = np.where(d > 0, n, 0).astype(np.uint16)
            H, b, cnt, sm, counts, costs = ref.odometry_coeffs(scale, true_rel, off, use_gradmag=gm)
            out[f"H{scale}{tag}"], out[f"b{scale}{tag}"] = H, b
            out[f"count{scale}{tag}"], out[f"sum{scale}{tag}"] = cnt, sm
            out[f"cost_counts{scale}{tag}"], out[f"cost_costs{scale}{tag}"] = counts, costs
        ref.close()
    os.makedirs("gpurun_out/golden", exist_ok=True)
    np.savez_compressed(f"gpurun_out/golden/{name}_odometry.npz", **out)
    print("wrote", name, "odometry: iterations", out["iterations"], "counts", [int(out[f"count{s}"]) for s in range(num_scales)],
          "error to the rendered motion", S.pose_error(out["est"], true_rel))


if __name__ == "__main__":
    if "--odometry-only" in sys.argv:
        golden_odometry("tiny")
        sys.exit(0)
    if "--preprocess-only" in sys.argv:
        golden_preprocess("tiny")
        sys.exit(0)
    if "--extra-only" in sys.argv:
        golden_intrinsics_pcg("tiny")
        golden_end_tasks("tiny")
        sys.exit(0)
    if "--end-tasks-only" in sys.argv:
        golden_end_tasks("tiny")
        sys.exit(0)
    golden_for("cfg1")
    golden_for("tiny")
    golden_for("tiny", True, False, "_depth_only")
    golden_for("tiny", False, True, "_desc_only")
    golden_intrinsics_pcg("tiny")
    golden_end_tasks("tiny")
    golden_preprocess("tiny")
    golden_odometry("tiny")
