"""GPU timing of one PCG outer iteration (ours vs the reference's kernels) on a BASELINE config."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from badslam_b200 import scene as S
from badslam_b200.direct_ba import DirectBA
from oracle import ref_cuda

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
inner_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = S.make_scene(S.config_by_name(name))
ba = DirectBA.from_scene(sc)
for rep in range(2):
    poses, act = sc.poses_init.copy(), np.zeros(sc.cfg.num_keyframes, np.int32)
    ba.SetKeyframeStates(poses, act)
    ba.SetSurfelsHost(sc.surfels, sc.num_surfels) if hasattr(ba, "SetSurfelsHost") else None
    t0 = time.time()
    r = ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, use_pcg=True, pcg_gauge_keyframe=0)
    t1 = time.time()
    print(f"ours {name} rep {rep}: wall {1e3*(t1-t0):.1f} ms, pcg {r.ms_pcg:.1f} ms, normals {r.ms_geometry_optimization:.2f} ms, inner {r.pcg_inner_iterations_total},"
          f" {r.ms_pcg / max(r.pcg_inner_iterations_total, 1):.2f} ms/inner, launches {r.kernel_launches}, r_norm {r.pcg_last_r_norm:.3f}", flush=True)
ref = ref_cuda.RefDirectBA(sc)
t0 = time.time()
rr = ref.bundle_adjust_pcg(min_iterations=1, max_iterations=1, max_inner_iterations=inner_ref, gauge_keyframe=0)
t1 = time.time()
print(f"ref  {name}: wall {1e3*(t1-t0):.1f} ms, pcg {rr.ms_pcg:.1f} ms, inner {rr.inner_iterations_total}, {rr.ms_pcg / rr.inner_iterations_total:.2f} ms/inner (incl. init), launches {rr.kernel_launches}")
