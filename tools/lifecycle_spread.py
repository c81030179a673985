"""Run-to-run spread of the reference's surfel creation / merging (its outcome depends on an atomicCAS race for sparse cell size > 1,
kernel_create_surfels.cu:68, kernel_supporting_surfels.cu) next to this backend's deterministic outcome: the data behind the
tolerances of tests/test_gpu_lifecycle.py.  Prints per keyframe: ours | several independent runs of the reference's kernels.

    python tools/lifecycle_spread.py [--runs 5]
"""
import argparse
import copy
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def half_map(S, name):
    sc = copy.copy(S.make_scene(S.config_by_name(name)))
    sc.poses_init = sc.poses_true.copy()
    sc.num_surfels = sc.num_surfels // 2
    cells = sc.cfactor.size * sc.cfg.num_keyframes
    sc.surfels = np.pad(sc.surfels, ((0, 0), (0, (cells + 127) // 128 * 128)))
    return sc


def main():
    from badslam_b200 import scene as S
    from badslam_b200.direct_ba import DirectBA
    from oracle import ref_cuda
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=5)
    a = ap.parse_args()
    for name in ("tiny", "small"):
        sc = half_map(S, name)
        K = sc.cfg.num_keyframes
        ba = DirectBA.from_scene(sc)
        refs = [ref_cuda.RefDirectBA(sc) for _ in range(a.runs)]
        print(f"== {name}: creation (filter_new_surfels = true), cell size {sc.cfg.cell}")
        for k in range(K):
            c0 = ba.CreateSurfelsForKeyframe(None, True, k)
            cr = [r.create_surfels_for_keyframe(k, True) for r in refs]
            print(f"keyframe {k}: ours {c0} | reference runs {cr} mean {np.mean(cr):.1f} std {np.std(cr, ddof=1):.1f} "
                  f"-> (ours - mean) / std = {(c0 - np.mean(cr)) / max(np.std(cr, ddof=1), 1e-9):.1f}, relative {(c0 - np.mean(cr)) / np.mean(cr):+.3f}")
        for r in refs:
            r.close()
        # merging on IDENTICAL surfels (tests/test_gpu_lifecycle.py::test_merge_surfels_three_way)
        seed = DirectBA.from_scene(sc)
        for k in range(K):
            seed.CreateSurfelsForKeyframe(None, False, k)
        rows = seed.GetSurfelsHost()
        sc2 = copy.copy(sc)
        sc2.surfels = sc.surfels.copy()
        sc2.num_surfels = rows.shape[1]
        sc2.surfels[:8, :sc2.num_surfels] = rows
        ba2 = DirectBA.from_scene(sc2)
        refs = [ref_cuda.RefDirectBA(sc2) for _ in range(a.runs)]
        tot0, totr = 0, [0] * a.runs
        for k in range(K):
            tot0 += ba2.MergeSurfelsForKeyframe(k)
            for i, r in enumerate(refs):
                totr[i] += r.merge_surfels_for_keyframe(k)
        print(f"== {name}: merged of {sc2.num_surfels}: ours {tot0} | reference runs {totr} mean {np.mean(totr):.1f} std {np.std(totr, ddof=1):.1f} "
              f"relative {(tot0 - np.mean(totr)) / np.mean(totr):+.3f}")
        for r in refs:
            r.close()


if __name__ == "__main__":
    main()
