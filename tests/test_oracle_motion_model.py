"""The constant-motion model in front of the image-pair odometry (BadSlam::PredictFramePose and the list handling around it,
bad_slam.cc:542-565, 767-827, 949-954, 1057-1068): the library's host functions (bba_host_motion_model_*) against
oracle/motion_model_oracle.py, and both against what the model must do by construction.  CPU only."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from badslam_b200 import _lib
from badslam_b200.direct_ba import MotionModel
from oracle import cpu_oracle as O
from oracle import motion_model_oracle as MO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_pose(rng, rot=0.2, trans=0.3):
    return O.se3_exp(np.concatenate([rng.uniform(-trans, trans, 3), rng.uniform(-rot, rot, 3)]).astype(np.float32))


def close(a, b, tol=2e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if np.dot(a[:4], b[:4]) < 0:   # q and -q are the same rotation
        b = np.concatenate([-b[:4], b[4:]])
    return np.abs(a - b).max() <= tol


def test_record_layout_matches_the_header():
    assert C.sizeof(_lib.MotionModelRecord) == 4 + 2 * 3 * 7 * 4
    assert _lib.MotionModelRecord.base_kf_tr_frame.offset == 4 and _lib.MotionModelRecord.frame_tr_base_kf.offset == 4 + 84


@pytest.mark.parametrize("use_motion_model", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_histories_match_the_oracle(seed, use_motion_model):
    """Any interleaving of push / rebase / clear: the stored lists and both predictions agree with the restatement after every
    operation (the two sides share no code: host_math.hpp vs oracle/badba_oracle.c)."""
    rng = np.random.default_rng(seed)
    ours, orc = MotionModel(use_motion_model), MO.MotionModel(use_motion_model)
    for step in range(60):
        op = rng.choice(["push", "push", "push", "rebase", "clear", "clear_identity"], p=[0.3, 0.25, 0.25, 0.1, 0.05, 0.05])
        if op == "push":
            e = random_pose(rng)
            ours.Push(e)
            orc.push(e)
        elif op == "rebase":
            ours.Rebase()
            orc.rebase()
        elif op == "clear":
            a, b = random_pose(rng, 1.0, 2.0), random_pose(rng, 1.0, 2.0)
            ours.Clear(a, b)
            orc.clear(a, b)
        else:
            ours.Clear()
            orc.clear()
        assert len(orc.base_kf_tr_frame) == ours.base_kf_tr_frame.shape[0] <= 3
        for mine, theirs in zip(ours.base_kf_tr_frame, orc.base_kf_tr_frame):
            assert close(mine, theirs), (step, op)
        for mine, theirs in zip(ours.frame_tr_base_kf, orc.frame_tr_base_kf):
            assert close(mine, theirs), (step, op)
        e1, e2 = ours.PredictFramePose()
        o1, o2 = orc.predict()
        assert close(e1, o1, 5e-6) and close(e2, o2, 5e-6), (step, op)


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_constant_twist_is_extrapolated(impl):
    """A camera moving with a constant twist relative to the base keyframe: after k frames both estimates are the pose of frame
    k + 1 (estimate 1 from the last two frames, estimate 2 from the two before the last)."""
    twist = np.array([0.02, -0.01, 0.03, 0.01, 0.02, -0.015], np.float32)
    step = O.se3_exp(twist)
    m = MotionModel() if impl == "product" else MO.MotionModel()
    push, predict = (m.Push, m.PredictFramePose) if impl == "product" else (m.push, m.predict)
    pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    for k in range(1, 7):
        pose = O.se3_mul(pose, step)   # frame k relative to the base keyframe
        push(pose)
        e1, e2 = predict()
        nxt = O.se3_mul(pose, step)
        assert close(e1, nxt, 1e-5), k
        if k >= 2:
            assert close(e2, nxt, 1e-5), k
        else:
            assert close(e2, e1, 0)   # fewer than three stored transforms: the second estimate is the first


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_second_estimate_ignores_the_last_frame(impl):
    """Estimate 2 exists for robustness against an outlier in the frame tracked last: it must not depend on it at all."""
    rng = np.random.default_rng(5)
    a, b = random_pose(rng), random_pose(rng)
    seconds = []
    for trial in range(3):
        m = MotionModel() if impl == "product" else MO.MotionModel()
        push, predict = (m.Push, m.PredictFramePose) if impl == "product" else (m.push, m.predict)
        push(a)
        push(b)
        push(random_pose(np.random.default_rng(100 + trial), 2.0, 5.0))   # the outlier
        seconds.append(predict()[1])
    assert np.array_equal(seconds[0], seconds[1]) and np.array_equal(seconds[0], seconds[2])
    # ... and it is b * (a^-1 b)^2: the step from a to b applied twice more
    d = O.se3_mul(O.se3_inverse(a), b)
    assert close(seconds[0], O.se3_mul(O.se3_mul(b, d), d), 5e-6)


@pytest.mark.parametrize("impl", ["product", "oracle"])
def test_rebase_commutes_with_prediction(impl):
    """Creating a keyframe from the last tracked frame changes the base the lists are expressed in, not the motion they describe:
    old_base_T_prediction == old_base_T_new_base * (prediction after the rebase)."""
    rng = np.random.default_rng(9)
    m = MotionModel() if impl == "product" else MO.MotionModel()
    push, predict, rebase = (m.Push, m.PredictFramePose, m.Rebase) if impl == "product" else (m.push, m.predict, m.rebase)
    last = None
    for _ in range(3):
        last = random_pose(rng, 0.05, 0.1)
        push(last)
    before = predict()
    rebase()
    after = predict()
    lists = m.base_kf_tr_frame
    assert close(np.asarray(lists[-1]), [0, 0, 0, 1, 0, 0, 0], 0)
    for e_before, e_after in zip(before, after):
        assert close(e_before, O.se3_mul(last, e_after), 1e-5)


def test_without_motion_model_and_edge_cases():
    m = MotionModel(use_motion_model=False)
    p = O.se3_exp(np.array([0.1, 0.2, 0.3, 0.05, 0.0, -0.05], np.float32))
    m.Push(p)
    m.Push(O.se3_mul(p, p))
    e1, e2 = m.PredictFramePose()
    assert np.array_equal(e1, O.se3_mul(p, p)) and np.array_equal(e1, e2)   # the last transform, untouched
    # an empty record predicts nothing; rebasing it makes it a one-entry identity history (bad_slam.cc:1062-1064)
    lib = _lib.load()
    rec = _lib.MotionModelRecord()
    out1, out2 = np.full(7, 7, np.float32), np.full(7, 7, np.float32)
    assert lib.bba_host_motion_model_predict(C.byref(rec), 1, out1.ctypes.data, out2.ctypes.data) == 0
    assert (out1 == 7).all() and (out2 == 7).all()
    lib.bba_host_motion_model_rebase(C.byref(rec))
    assert rec.count == 1 and list(rec.base_kf_tr_frame[0]) == [0, 0, 0, 1, 0, 0, 0]
    # the history never grows beyond three transforms and drops the oldest first
    m = MotionModel()
    poses = [O.se3_exp(np.array([0.01 * k, 0, 0, 0, 0, 0], np.float32)) for k in range(1, 6)]
    for q in poses:
        m.Push(q)
    assert np.array_equal(m.base_kf_tr_frame, np.stack(poses[-3:]))
    # null arguments are ignored, not dereferenced
    lib.bba_host_motion_model_push(None, None)
    lib.bba_host_motion_model_rebase(None)
    lib.bba_host_motion_model_clear(None, None, None)
    assert lib.bba_host_motion_model_predict(None, 1, out1.ctypes.data, out2.ctypes.data) == 0


def test_cpp_adaptor_motion_model(tmp_path):
    """include/badba_direct_ba.hpp's MotionModel<SE3f> keeps the reference's method names; compile it with a plain pose type and
    run the RunOdometry sequence (predict, push) and the keyframe rebase on the host."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no C++ compiler")
    src = tmp_path / "mm.cpp"
    src.write_text(r'''
#include <cmath>
#include "badba_direct_ba.hpp"
struct SE3f { float v[7] = {0, 0, 0, 1, 0, 0, 0}; float* data() { return v; } const float* data() const { return v; } };
int main() {
  badba::MotionModel<SE3f> mm(true);
  SE3f e1, e2, est;
  for (int k = 1; k <= 4; ++k) {
    mm.PredictFramePose(&e1, &e2);
    if (std::fabs(e1.v[4] - 0.1f * k) > 1e-6f && k > 2) return 1;   // constant velocity along x from the third frame on
    est.v[4] = 0.1f * k;                                           // "tracking result"
    mm.Push(est);
  }
  if (mm.stored_frames() != 3) return 2;
  mm.Rebase();
  mm.PredictFramePose(&e1, &e2);
  if (std::fabs(e1.v[4] - 0.1f) > 1e-6f || std::fabs(e2.v[4] - 0.1f) > 1e-6f) return 3;   // one more step, seen from the new keyframe
  SE3f kf, frame;
  kf.v[4] = -1.f; frame.v[4] = 1.25f;
  mm.ClearMotionModel(&kf, &frame);
  mm.PredictFramePose(&e1, &e2);
  return (mm.stored_frames() == 1 && std::fabs(e1.v[4] - 0.25f) < 1e-6f) ? 0 : 4;
}''')
    exe = tmp_path / "mm"
    libdir = os.path.join(ROOT, "badslam_b200")
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", str(src), "-o",
                           str(exe), "-L", libdir, "-lbadba_b200", f"-Wl,-rpath,{libdir}"])
    assert subprocess.call([str(exe)]) == 0
