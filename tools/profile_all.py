"""Driver for ONE ncu capture of every hot kernel (profiles/): run under
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'PoseAccumulate|ActivationNormals|PositionDescriptor|IntrinsicsAccumulate|PcgAccumulate|ObservationStats' ...
The profiled range holds one alternating-BA iteration, one intrinsics step, one PCG init + sweep and the end tasks."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from badslam_b200.direct_ba import DirectBA
from badslam_b200.scene import config_by_name, make_scene

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
sc = make_scene(config_by_name(name))
ba = DirectBA.from_scene(sc)
surf = ba.surfels()
backup = surf[:8].clone()
act0 = np.zeros(sc.cfg.num_keyframes, np.int32)
ba.SetLastBAIterationCount(ba.ba_iteration_count())


def reset():
    surf[:8].copy_(backup)
    ba.SetKeyframeStates(sc.poses_init, act0)


for _ in range(2):   # warm-up (not profiled)
    reset()
    ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)
reset()
torch.cuda.synchronize()
torch.cuda.profiler.start()
r = ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)
ba.OptimizeIntrinsics(True, True)
reset()
ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, use_pcg=True, pcg_max_inner_iterations=1, pcg_gauge_keyframe=0,
                    increase_ba_iteration_count=False)
reset()
ba.PerformBASchemeEndTasks()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", r.ms_pose_optimization, ba.surfels_size())
