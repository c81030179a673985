"""Minimal driver for ncu: builds a scene and runs a few outer BA iterations (same step as bench.py)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from badslam_b200.direct_ba import DirectBA
from badslam_b200.scene import config_by_name, make_scene

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = make_scene(config_by_name(name))
ba = DirectBA.from_scene(sc)
surf = ba.surfels()
backup = surf[:8].clone()
act0 = np.zeros(sc.cfg.num_keyframes, np.int32)
for i in range(steps):
    surf[:8].copy_(backup)
    ba.SetKeyframeStates(sc.poses_init, act0)
    r = ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)
torch.cuda.synchronize()
print("done", r.ms_pose_optimization)
