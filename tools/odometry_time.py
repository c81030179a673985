"""Times the image-pair odometry (bba_track_frame_pairwise: one brightness launch, one level-0 launch, one launch per pyramid
level, ONE persistent coarse-to-fine Gauss-Newton kernel) against the reference's own kernels behind its host loop
(oracle/_ref: per iteration 2-4 clears + kernel + 2 device-to-host copies + stream synchronisation) on the same frame pair.

    python tools/odometry_time.py [--size 640x480] [--scales 5] [--reps 20]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from badslam_b200 import scene as S
    from badslam_b200.direct_ba import DirectBA
    from oracle import ref_cuda
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="640x480")
    ap.add_argument("--scales", type=int, default=5)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    w, h = [int(v) for v in a.size.split("x")]
    sc = S.make_scene(S.SceneConfig(width=w, height=h, num_keyframes=2, num_surfels=2000, cell=4, seed=31, name="odometry"))
    true_rel = S.se3_exp([0.02, -0.01, 0.015, 0.01, -0.008, 0.012])
    depth, normals, _, color = S.render_frame(sc, S.se3_mul(sc.poses_true[0], true_rel))
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    init2 = S.se3_exp([0.01, 0, 0, 0, 0, 0])
    ba, ref = DirectBA.from_scene(sc), ref_cuda.RefDirectBA(sc)
    dev = (torch.from_numpy(depth.view(np.int16)).cuda(), torch.from_numpy(normals.view(np.int16)).cuda(), torch.from_numpy(color).cuda())
    for _ in range(3):
        est, res = ba.TrackFramePairwise(None, 0, *dev, ident, init2, num_scales=a.scales)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.reps):
        est, res = ba.TrackFramePairwise(None, 0, *dev, ident, init2, num_scales=a.scales)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.reps * 1e3
    ours = e0.elapsed_time(e1) / a.reps
    ref_ms = []
    for _ in range(4):
        est_r, res_r = ref.track_frame_pairwise(0, depth, normals, color, ident, init2, num_scales=a.scales)
        ref_ms.append(res_r.ms)
    its, its_r = list(res.iterations)[:a.scales], list(res_r.iterations)[:a.scales]
    print(f"image-pair odometry {w}x{h}, {a.scales} pyramid levels: {ours:.3f} ms per frame on the device ({wall:.3f} ms wall incl. the result "
          f"copy), {res.kernel_launches} launches, {res.passes} image passes, Gauss-Newton iterations per level {its}")
    print(f"reference kernels + host loop: {min(ref_ms[1:]):.3f} ms per frame, {res_r.kernel_launches} launches, iterations {its_r}  "
          f"-> {min(ref_ms[1:]) / ours:.1f}x")
    print(f"error to the rendered motion: ours {S.pose_error(est, true_rel)}, reference {S.pose_error(est_r, true_rel)}, "
          f"difference {S.pose_error(est, est_r)}")


if __name__ == "__main__":
    main()
