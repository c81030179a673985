"""GPU exploration: PCG-based BA, ours vs the reference's CUDA kernels vs the CPU oracle."""
import dataclasses, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from badslam_b200 import scene as S
from badslam_b200.direct_ba import DirectBA
from oracle import cpu_oracle as O, ref_cuda


def perr(a, b, K):
    return max(max(S.pose_error(a[k], b[k])) for k in range(K))


def run(name, distort, intr, outer=2, with_oracle=True, use_depth=True, use_desc=True):
    cfg = S.config_by_name(name)
    if distort:
        cfg = dataclasses.replace(cfg, depth_a=0.03, cfactor=0.005)
    sc = S.make_scene(cfg)
    K = cfg.num_keyframes
    ba = DirectBA.from_scene(sc, use_depth_residuals=use_depth, use_descriptor_residuals=use_desc)
    ref = ref_cuda.RefDirectBA(sc, use_depth, use_desc)
    ref2 = ref_cuda.RefDirectBA(sc, use_depth, use_desc)
    orc = O.Oracle(sc, use_depth, use_desc) if with_oracle else None
    for it in range(outer):
        t0 = time.time()
        r0 = ba.BundleAdjustment(None, intr, intr, False, True, True, 1, 1, use_pcg=True, pcg_gauge_keyframe=1)
        t1 = time.time()
        r1 = ref.bundle_adjust_pcg(True, True, intr, intr, 1, 1, 30, 1)
        t2 = time.time()
        r2 = ref2.bundle_adjust_pcg(True, True, intr, intr, 1, 1, 30, 1)
        p0, _ = ba.GetKeyframeStates(); p1 = ref.poses(); p2 = ref2.poses()
        s0, s1, s2 = ba.GetSurfelsHost(), ref.surfels(), ref2.surfels()
        print(f"--- {name} distort={distort} intr={intr} depth={use_depth} desc={use_desc} outer {it}: ours {1e3*(t1-t0):.1f} ms (pcg {r0.ms_pcg:.2f}) ref {1e3*(t2-t1):.1f} ms (pcg {r1.ms_pcg:.2f})"
              f" launches {r0.kernel_launches}/{r1.kernel_launches}")
        print(f"  inner ours {r0.pcg_inner_iterations_total} ref {r1.inner_iterations_total} ref2 {r2.inner_iterations_total}; r_norm {r0.pcg_last_r_norm:.6f} {r1.last_r_norm:.6f} {r2.last_r_norm:.6f}")
        print(f"  pose ours-ref {perr(p0, p1, K):.2e} ref-ref2 {perr(p1, p2, K):.2e}; surfel pos max {np.abs(s0[:3]-s1[:3]).max():.2e} (ref-ref2 {np.abs(s1[:3]-s2[:3]).max():.2e})"
              f" desc max {np.abs(s0[6:8]-s1[6:8]).max():.2e} (ref-ref2 {np.abs(s1[6:8]-s2[6:8]).max():.2e}) normals differ {(s0[3].view(np.uint32)!=s1[3].view(np.uint32)).sum()}")
        if intr:
            d0, c0, a0 = ba._intrinsics(); d1, c1, a1 = ref.intrinsics(); d2, c2, a2 = ref2.intrinsics()
            print(f"  depth_K ours-ref {np.abs(d0-d1).max():.2e} (ref-ref2 {np.abs(d1-d2).max():.2e}) color_K {np.abs(c0-c1).max():.2e} ({np.abs(c1-c2).max():.2e}) a {a0:.6f} {a1:.6f} {a2:.6f}"
                  f" cf {np.abs(ba.cfactor_buffer()-ref.cfactor()).max():.2e} ({np.abs(ref.cfactor()-ref2.cfactor()).max():.2e})")
        if orc is not None:
            ro = orc.bundle_adjust_pcg(True, True, intr, intr, 1, 1, 30, 1)
            print(f"  oracle inner {ro.inner_iterations_total} r_norm {ro.last_r_norm:.6f} pose ours-orc {perr(p0, orc.poses, K):.2e} surfel pos {np.abs(s0[:3]-orc.surfels[:3,:sc.num_surfels]).max():.2e}")


if __name__ == "__main__":
    run("tiny", False, False)
    run("tiny", False, False, use_desc=False)
    run("small", True, True)
    run("cfg2", False, False, outer=1, with_oracle=False)
