"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the constant-motion model BadSlam keeps in front of its image-pair odometry.

What it follows (applications/badslam/src/badslam/bad_slam.cc):
  ClearMotionModel    :542-565    one stored transform: last keyframe's frame_T_global * the frame's global_T_frame (or identity)
  PredictFramePose    :767-827    estimate 1 = last * inverse(previous) * last; estimate 2 = previous * step * step with
                                  step = inverse(before previous) * previous; fall-backs with fewer than 2 / 3 stored transforms;
                                  without the motion model both are the last transform
  RunOdometry, tail   :949-954    at most three transforms are kept, the new estimate and its inverse are appended
  ProcessFrame        :1057-1068  after a keyframe was created from the last tracked frame: older entries are re-expressed
                                  relative to it (inverse list: F_i * B_last, forward list: F_last * B_i), the last becomes identity

The transforms are Sophus::SE3f products; they are taken from the oracle's C restatement of Sophus (orc_se3_mul / orc_se3_inverse,
oracle/badba_oracle.c), evaluated left to right like the reference's `a * b * c`.  Two lists (transform and inverse) are carried
side by side exactly as the reference does; it never re-derives one from the other.

Pinning: PARITY UNPINNED against the reference binary -- BadSlam (the class that owns this state) needs the whole application to
build and has no test or golden vector for the motion model.  The restatement is anchored on what the model must do by
construction (tests/test_oracle_motion_model.py): a constant twist is extrapolated exactly, estimate 2 ignores the last frame,
rebasing commutes with prediction.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

from . import cpu_oracle as O

IDENTITY = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)


class MotionModel:
    def __init__(self, use_motion_model: bool = True):
        self.use_motion_model = use_motion_model
        self.base_kf_tr_frame = [IDENTITY.copy()]
        self.frame_tr_base_kf = [IDENTITY.copy()]

    def clear(self, last_kf_frame_T_global=None, global_T_frame=None):   # bad_slam.cc:542-565
        if last_kf_frame_T_global is None:
            self.base_kf_tr_frame, self.frame_tr_base_kf = [IDENTITY.copy()], [IDENTITY.copy()]
            return
        rel = O.se3_mul(np.asarray(last_kf_frame_T_global, np.float32), np.asarray(global_T_frame, np.float32))
        self.base_kf_tr_frame, self.frame_tr_base_kf = [rel], [O.se3_inverse(rel)]

    def predict(self):   # bad_slam.cc:767-827
        B, F = self.base_kf_tr_frame, self.frame_tr_base_kf
        n = len(B)
        if not self.use_motion_model:
            return B[n - 1].copy(), B[n - 1].copy()
        e1 = O.se3_mul(O.se3_mul(B[n - 1], F[n - 2]), B[n - 1]) if n >= 2 else B[n - 1].copy()
        if n >= 3:
            step = O.se3_mul(F[n - 3], B[n - 2])
            e2 = O.se3_mul(O.se3_mul(B[n - 2], step), step)
        else:
            e2 = e1.copy()
        return e1, e2

    def push(self, base_T_frame_estimate):   # bad_slam.cc:949-954
        e = np.asarray(base_T_frame_estimate, np.float32).copy()
        if len(self.base_kf_tr_frame) >= 3:
            del self.base_kf_tr_frame[0]
            del self.frame_tr_base_kf[0]
        self.base_kf_tr_frame.append(e)
        self.frame_tr_base_kf.append(O.se3_inverse(e))

    def rebase(self):   # bad_slam.cc:1057-1068
        B, F = self.base_kf_tr_frame, self.frame_tr_base_kf
        for i in range(len(F) - 1):
            F[i] = O.se3_mul(F[i], B[-1])
            B[i] = O.se3_mul(F[-1], B[i])
        if not F:
            B.append(IDENTITY.copy())
            F.append(IDENTITY.copy())
        else:
            B[-1], F[-1] = IDENTITY.copy(), IDENTITY.copy()
