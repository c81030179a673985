"""Locality A/B for the keyframe-pixel gathers (DESIGN.md "TMA tiles vs gathers").

north_star asks for keyframe pixels staged through TMA; SURVEY 7.4 proposes binning the surfels by (keyframe, image tile) so that
a 2-D cp.async.bulk.tensor tile can be staged per bin.  Whatever such a scheme costs (binning passes, the staging itself), the
BEST it can do is to make every pixel read perfectly local.  This tool measures that bound without building the scheme:

  workload A  K keyframes that are byte-identical COPIES of keyframe 0 (same pose, own memory, own luma array): the kernels run
              exactly the instruction stream of workload B, but gather from K x 1.5 MB of distinct memory like the real scene;
  workload B  the same K keyframes ALIASED to keyframe 0's buffers and luma array: all gathers of all keyframes hit the same
              1.5 MB, L1 / L2 resident -- perfect locality, nothing else changed (same associations, same residual counts).

    python tools/ab_locality.py [--workload cfg3] [--steps 5]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(workload, steps, alias):
    import numpy as np
    import torch
    from badslam_b200 import scene as S
    from badslam_b200.direct_ba import DirectBA, Keyframe, PinholeCamera4f
    sc = S.make_scene(S.config_by_name(workload))
    cfg = sc.cfg
    K = cfg.num_keyframes
    dev = torch.device("cuda", 0)
    cam = PinholeCamera4f(cfg.width, cfg.height, sc.depth_K)
    ba = DirectBA(max_surfel_count=sc.pitch, raw_to_float_depth=cfg.raw_to_float_depth, baseline_fx=cfg.baseline_fx,
                  sparse_surfel_cell_size=cfg.cell, color_camera_initial_estimate=cam, depth_camera_initial_estimate=cam,
                  device=dev, max_keyframes=K)
    first = Keyframe.from_host(0, sc.depth[0], sc.normals[0], sc.radius[0], sc.color[0], sc.poses_init[0], sc.min_depth[0], sc.max_depth[0], dev)
    ba.AddKeyframe(first)
    for k in range(1, K):
        kf = Keyframe.__new__(Keyframe)
        kf.__dict__.update(first.__dict__)
        kf.frame_index = k
        if not alias:
            kf.depth_buffer, kf.normals_buffer = first.depth_buffer.clone(), first.normals_buffer.clone()
            kf.radius_buffer, kf.color_buffer = first.radius_buffer.clone(), first.color_buffer.clone()
        ba.AddKeyframe(kf)
    surf = torch.from_numpy(sc.surfels).to(dev)
    ba.SetSurfels(surf, sc.num_surfels)
    backup = surf[:8].clone()
    poses0 = np.repeat(sc.poses_init[:1], K, axis=0)
    act0 = np.zeros(K, np.int32)
    ba.SetLastBAIterationCount(ba.ba_iteration_count())

    def step():
        surf[:8].copy_(backup, non_blocking=True)
        ba.SetKeyframeStates(poses0, act0)
        return ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)

    for _ in range(3):
        res = step()
    torch.cuda.synchronize()
    ba.SetProfiling(1)
    ba.GetProfile(reset=True)
    stage = np.zeros(3)
    for _ in range(steps):
        res = step()
        stage += [res.ms_surfel_activation, res.ms_geometry_optimization, res.ms_pose_optimization]
    torch.cuda.synchronize()
    prof = ba.GetProfile(reset=True)
    print("AB_RESULT " + json.dumps({
        "aliased": bool(alias), "stage_ms": [round(float(v) / steps, 3) for v in stage],
        "pose_kernel_avg_ms": round(prof["pose_ms"] / max(prof["pose_launches"], 1), 4),
        "pose_launches_per_step": prof["pose_launches"] / steps,
        "residuals": int(res.depth_residual_count + res.descriptor_residual_count), "gn_iterations": int(res.pose_iterations_total)}), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        return child(args[1], int(args[2]), args[3] == "1")
    workload, steps = "cfg3", 5
    while args:
        if args[0] == "--workload":
            workload = args[1]
        elif args[0] == "--steps":
            steps = int(args[1])
        args = args[2:]
    for alias in (0, 1):
        env = dict(os.environ)
        if alias:
            env["BADBA_ALIAS_LUMA"] = "1"
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", workload, str(steps), str(alias)], env=env,
                           capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("AB_RESULT ")]
        print(("B aliased (perfect locality): " if alias else "A distinct copies (real locality): ") +
              (line[-1][10:] if line else f"FAILED rc={p.returncode} {p.stderr[-800:]}"), flush=True)


if __name__ == "__main__":
    main()
