"""GPU exploration: OptimizeIntrinsics with a non-zero depth deformation model (a, cfactor)."""
import dataclasses, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from badslam_b200 import scene as S
from badslam_b200.direct_ba import DirectBA
from oracle import cpu_oracle, ref_cuda

cfg = dataclasses.replace(S.config_by_name("small"), depth_a=0.03, cfactor=0.005)
sc = S.make_scene(cfg)
rng = np.random.default_rng(5)
cf_init = (rng.standard_normal(sc.cfactor.shape) * 0.003).astype(np.float32)
a_init = 0.02
for rep in range(2):
    ba, ref, orc = DirectBA.from_scene(sc), ref_cuda.RefDirectBA(sc), cpu_oracle.Oracle(sc)
    ba.SetA(a_init); ba.SetCFactorBuffer(cf_init)
    ref.set_depth_params(a_init, cf_init)
    orc.model.a = a_init; orc.cfactor[:] = cf_init
    for step in range(3):
        ba.OptimizeIntrinsics(True, True); ref.optimize_intrinsics(True, True); orc.optimize_intrinsics(True, True)
        d0, c0, a0 = ba._intrinsics(); d1, c1, a1 = ref.intrinsics()
        d2 = np.array(orc.model.depth_K[:], np.float32); a2 = orc.model.a
        print(f"rep {rep} step {step}: ours a {a0:.6f} ref a {a1:.6f} orc a {a2:.6f} | dK ours-ref {np.abs(d0-d1).max():.2e} ours-orc {np.abs(d0-d2).max():.2e}"
              f" | cf ours-ref {np.abs(ba.cfactor_buffer()-ref.cfactor()).max():.2e} ours-orc {np.abs(ba.cfactor_buffer()-orc.cfactor).max():.2e} | K {d0}")
