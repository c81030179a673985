"""GPU exploration: one OptimizeIntrinsics step, ours vs the reference's CUDA kernels vs the CPU oracle."""
import dataclasses
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from badslam_b200 import scene as S  # noqa: E402
from badslam_b200.direct_ba import DirectBA, PinholeCamera4f  # noqa: E402
from oracle import cpu_oracle, ref_cuda  # noqa: E402


def run(name, distort, perturb, with_oracle=True):
    cfg = S.config_by_name(name)
    if distort:
        cfg = dataclasses.replace(cfg, depth_a=0.03, cfactor=0.005)
    sc = S.make_scene(cfg)
    if perturb:
        sc.depth_K = (np.asarray(sc.depth_K, np.float32) * np.float32([1.003, 0.998, 1.002, 0.997])).astype(np.float32)
        sc.color_K = (np.asarray(sc.color_K, np.float32) * np.float32([0.998, 1.002, 1.001, 0.999])).astype(np.float32)
    for od, oc in ((True, True), (True, False), (False, True)):
        ba, ref = DirectBA.from_scene(sc), ref_cuda.RefDirectBA(sc)
        t0 = time.time(); ba.OptimizeIntrinsics(od, oc); t1 = time.time()
        ref.optimize_intrinsics(od, oc); t2 = time.time()
        d0, c0, a0 = ba._intrinsics()
        d1, c1, a1 = ref.intrinsics()
        cf0, cf1 = ba.cfactor_buffer(), ref.cfactor()
        print(f"--- {name} distort={distort} perturb={perturb} depth={od} color={oc}: ours {1e3*(t1-t0):.1f} ms ref {1e3*(t2-t1):.1f} ms")
        print("  init  depth_K", np.asarray(sc.depth_K), "color_K", np.asarray(sc.color_K))
        print("  ours  depth_K", d0, "color_K", c0, "a", a0)
        print("  ref   depth_K", d1, "color_K", c1, "a", a1)
        print("  |d depth_K|", np.abs(d0 - d1), "|d color_K|", np.abs(c0 - c1), "|da|", abs(a0 - a1))
        print("  cfactor ours range", cf0.min(), cf0.max(), "max|d| vs ref", np.abs(cf0 - cf1).max(), "mean|d|", np.abs(cf0 - cf1).mean(),
              "nonzero", (cf0 != 0).sum(), (cf1 != 0).sum())
        if with_oracle:
            orc = cpu_oracle.Oracle(sc)
            orc.optimize_intrinsics(od, oc)
            d2 = np.array(orc.model.depth_K[:], np.float32); c2 = np.array(orc.model.color_K[:], np.float32); a2 = orc.model.a
            print("  orc   depth_K", d2, "color_K", c2, "a", a2)
            print("  |ours-orc| depth_K", np.abs(d0 - d2), "color_K", np.abs(c0 - c2), "a", abs(a0 - a2), "cf", np.abs(cf0 - orc.cfactor).max())
        # second run of ours for run-to-run noise
        ba2 = DirectBA.from_scene(sc); ba2.OptimizeIntrinsics(od, oc)
        d3, c3, a3 = ba2._intrinsics()
        print("  ours run-to-run depth_K", np.abs(d0 - d3), "color_K", np.abs(c0 - c3), "a", abs(a0 - a3), "cf", np.abs(cf0 - ba2.cfactor_buffer()).max())
        ref2 = ref_cuda.RefDirectBA(sc); ref2.optimize_intrinsics(od, oc)
        d4, c4, a4 = ref2.intrinsics()
        print("  ref  run-to-run depth_K", np.abs(d1 - d4), "color_K", np.abs(c1 - c4), "a", abs(a1 - a4), "cf", np.abs(cf1 - ref2.cfactor()).max())
    # full BA with intrinsics
    ba, ref = DirectBA.from_scene(sc), ref_cuda.RefDirectBA(sc)
    r = ba.BundleAdjustment(None, True, True, False, True, True, 3, 3, False, 0, cfg.num_keyframes - 1, True)
    rr = ref.bundle_adjust(min_iterations=3, max_iterations=3, optimize_depth_intrinsics=True, optimize_color_intrinsics=True)
    d0, c0, a0 = ba._intrinsics(); d1, c1, a1 = ref.intrinsics()
    p0, _ = ba.GetKeyframeStates(); p1 = ref.poses()
    errs = [S.pose_error(p0[k], p1[k]) for k in range(cfg.num_keyframes)]
    et, er = [e[0] for e in errs], [e[1] for e in errs]
    print(f"=== BA+intrinsics {name}: iters {r.iterations_done}/{rr.iterations_done} |d depth_K| {np.abs(d0-d1)} |d color_K| {np.abs(c0-c1)} |da| {abs(a0-a1)}"
          f" pose err t {np.max(et):.2e} r {np.max(er):.2e} cf {np.abs(ba.cfactor_buffer()-ref.cfactor()).max():.2e} ms_intr {r.ms_intrinsics_optimization:.3f}")
    print("   final depth_K", d0, "true", np.asarray(getattr(sc, 'depth_K_true', sc.depth_K)))


if __name__ == "__main__":
    run("tiny", False, False)
    run("small", True, True)
    run("cfg2", True, True, with_oracle=False)
