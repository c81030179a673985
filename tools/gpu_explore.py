"""First-contact GPU script: three-way comparison product CUDA vs CPU oracle vs reference CUDA.

Run on the GPU box:  python tools/gpu_explore.py [scene-name]
Writes a report to gpurun_out/explore_<scene>.txt.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from badslam_b200.direct_ba import DirectBA  # noqa: E402
from badslam_b200.scene import config_by_name, make_scene, pose_error  # noqa: E402
from oracle import cpu_oracle, ref_cuda  # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    os.makedirs("gpurun_out", exist_ok=True)
    out = open(f"gpurun_out/explore_{name}.txt", "w")

    def log(*a):
        s = " ".join(str(x) for x in a)
        print(s, flush=True)
        out.write(s + "\n")
        out.flush()

    log("device", torch.cuda.get_device_name(0))
    t = time.time()
    sc = make_scene(config_by_name(name))
    log(f"scene {name}: K={sc.cfg.num_keyframes} n={sc.num_surfels} gen {time.time() - t:.1f}s")
    K = sc.cfg.num_keyframes

    ba = DirectBA.from_scene(sc)
    ref = ref_cuda.RefDirectBA(sc)
    log("covis equal (product vs oracle):", end_covis(ba, sc))

    # --- which texture-weight emulation matches the hardware?
    for mode in (0, 1, 2):
        cpu_oracle.lib().orc_set_tex_mode(mode)
        orc = cpu_oracle.Oracle(sc)
        errs = []
        for k in range(min(K, 3)):
            st = orc.pose_coeffs(k)
            H, b, cnt, cost = ref.pose_coeffs(k, sc.poses_init[k])
            errs.append((rel(np.array(st.H[:]), H), rel(np.array(st.b[:]), b), int(st.n_assoc + st.n_photo) - cnt,
                         (st.cost_depth + st.cost_desc1 - cost) / cost))
        log(f"tex_mode {mode}: oracle vs REF  (relH, relb, dcount, relcost) = {errs}")
    cpu_oracle.lib().orc_set_tex_mode(1)
    orc = cpu_oracle.Oracle(sc)

    # --- pose coefficients, three way
    for k in range(min(K, 4)):
        st = orc.pose_coeffs(k)
        H, b, cnt, cost = ref.pose_coeffs(k, sc.poses_init[k])
        pc = ba.AccumulatePoseEstimationCoeffs(k, sc.poses_init[k])
        log(f"kf {k}: counts oracle {st.n_assoc}+{st.n_photo} ref {cnt} ours {pc.n_assoc}+{pc.n_photo} | "
            f"inimg {st.n_inimg}/{pc.n_inimg} depthok {st.n_depthok}/{pc.n_depthok}")
        log(f"   relH ours-ref {rel(pc.H[:], H):.2e} ours-oracle {rel(pc.H[:], st.H[:]):.2e} ref-oracle {rel(H, st.H[:]):.2e}")
        log(f"   relb ours-ref {rel(pc.b[:], b):.2e} ours-oracle {rel(pc.b[:], st.b[:]):.2e} ref-oracle {rel(b, st.b[:]):.2e}")
        log(f"   cost ours {pc.cost_depth + pc.cost_desc1:.6f} ref {cost:.6f} oracle {st.cost_depth + st.cost_desc1:.6f}")

    # --- EstimateFramePose
    for k in range(min(K, 4)):
        pr, ir, cr = ref.estimate_frame_pose(k, sc.poses_init[k])
        po, io, co = orc.estimate_frame_pose(k)
        pp, ip, cp = ba.EstimateFramePose(None, sc.poses_init[k], k)
        log(f"pose kf {k}: iters ref {ir} oracle {io} ours {ip}; ours-ref {pose_error(pp, pr)} oracle-ref {pose_error(po, pr)} "
            f"to-truth {pose_error(pp, sc.poses_true[k])}")

    # --- activation + geometry
    ref.update_activation()
    orc.update_activation()
    ba.UpdateSurfelActivation()
    a_ref, a_orc, a_our = ref.active(), orc.active[:sc.num_surfels], ba.GetActiveHost()
    log(f"activation: ref {a_ref.sum()} oracle {a_orc.sum()} ours {a_our.sum()} mismatches ours-ref {(a_ref != a_our).sum()} "
        f"oracle-ref {(a_ref != a_orc).sum()}")
    ref.optimize_geometry_iteration()
    orc.optimize_geometry_iteration()
    ba.OptimizeGeometryIteration()
    s_ref, s_orc, s_our = ref.surfels(), orc.surfels[:8, :sc.num_surfels], ba.GetSurfelsHost()
    for r, nm in ((0, "x"), (1, "y"), (2, "z"), (6, "d1"), (7, "d2")):
        log(f"geometry row {nm}: max|ours-ref| {np.max(np.abs(s_our[r] - s_ref[r])):.3e} max|oracle-ref| "
            f"{np.max(np.abs(s_orc[r] - s_ref[r])):.3e}  moved(max) {np.max(np.abs(s_ref[r] - sc.surfels[r, :sc.num_surfels])):.3e}")
    nb = lambda s: s[3].view(np.uint32)
    log(f"normals: packed mismatches ours-ref {(nb(s_our) != nb(s_ref)).sum()} oracle-ref {(nb(s_orc) != nb(s_ref)).sum()}")

    # --- full BA from fresh state
    ba2 = DirectBA.from_scene(sc)
    ref2 = ref_cuda.RefDirectBA(sc)
    orc2 = cpu_oracle.Oracle(sc)
    t0 = time.time(); r_our = ba2.BundleAdjustment(None, False, False, False, True, True, 10, 10); t_our = time.time() - t0
    t0 = time.time(); r_ref = ref2.bundle_adjust(True, True, 10, 10); t_ref = time.time() - t0
    t0 = time.time(); r_orc = orc2.bundle_adjust(True, True, 10, 10); t_orc = time.time() - t0
    log(f"BA ours: it {r_our.iterations_done} conv {r_our.converged} pose_its {r_our.pose_iterations_total} counts "
        f"{r_our.depth_residual_count}+{r_our.descriptor_residual_count} cost {r_our.cost:.4f} launches {r_our.kernel_launches} wall {t_our:.3f}s "
        f"stage ms {r_our.ms_surfel_activation:.3f} {r_our.ms_geometry_optimization:.3f} {r_our.ms_pose_optimization:.3f}")
    log(f"BA ref : it {r_ref.iterations_done} conv {r_ref.converged} pose_its {r_ref.pose_iterations_total} count {r_ref.n_count} cost {r_ref.cost:.4f} "
        f"launches {r_ref.kernel_launches} wall {t_ref:.3f}s stage ms {r_ref.ms_surfel_activation:.3f} {r_ref.ms_geometry_optimization:.3f} {r_ref.ms_pose_optimization:.3f}")
    log(f"BA orcl: it {r_orc.iterations_done} conv {r_orc.converged} pose_its {r_orc.pose_iterations_total} counts {r_orc.n_assoc}+{2 * r_orc.n_photo} cost {r_orc.cost:.4f} wall {t_orc:.3f}s")
    errs_or, errs_oo, errs_t = [], [], []
    for k in range(K):
        errs_or.append(pose_error(ba2.keyframes()[k].global_T_frame(), ref2.pose(k)))
        errs_oo.append(pose_error(orc2.poses[k], ref2.pose(k)))
        errs_t.append(pose_error(ba2.keyframes()[k].global_T_frame(), sc.poses_true[k]))
    log("BA final pose ours-ref max (m, rad):", np.max(np.array(errs_or), axis=0))
    log("BA final pose oracle-ref max (m, rad):", np.max(np.array(errs_oo), axis=0))
    log("BA final pose ours-truth max (m, rad):", np.max(np.array(errs_t), axis=0), " init-truth:",
        np.max(np.array([pose_error(sc.poses_init[k], sc.poses_true[k]) for k in range(K)]), axis=0))
    s_ref, s_our = ref2.surfels(), ba2.GetSurfelsHost()
    for r, nm in ((0, "x"), (1, "y"), (2, "z"), (6, "d1"), (7, "d2")):
        d = np.abs(s_our[r] - s_ref[r])
        log(f"BA final surfel row {nm}: max|ours-ref| {d.max():.3e} mean {d.mean():.3e}")
    out.close()


def end_covis(ba, sc):
    orc = cpu_oracle.Oracle(sc)
    return bool(np.array_equal(ba.covisibility(), orc.covis))


if __name__ == "__main__":
    main()
