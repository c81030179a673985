"""CPU-only: the oracle's PerformBASchemeEndTasks restatement (delete badly observed surfels, update radii, compact;
direct_ba.cc:566-653, kernel_delete_surfels.cu:42-164, kernel_compact_surfels.cu:159-279) through its properties."""
import numpy as np

from badslam_b200 import scene as S
from oracle import cpu_oracle as O

DELETED = np.uint32(0x7fffffff)


def perturbed_oracle():
    """tiny scene with surfels moved out of view, in front of the surface and behind it (scene.displace_surfels)."""
    sc, away, front, behind = S.displace_surfels(S.make_scene(S.config_by_name("tiny")))
    return sc, O.Oracle(sc), away, front, behind


def expected_compaction(rows, invalid):
    """kernel_compact_surfels.cu:126-157: the r-th valid surfel counted from the END moves into the r-th free spot."""
    n = len(invalid)
    free = np.flatnonzero(invalid)
    valid_desc = np.flatnonzero(~invalid)[::-1]
    out = rows.copy()
    for r, src in enumerate(valid_desc[:len(free)]):
        if free[r] < src:
            out[:, free[r]] = rows[:, src]
    return out[:, :n - len(free)]


def test_end_tasks_delete_update_radii_and_compact():
    sc, orc, away, front, behind = perturbed_oracle()
    n = sc.num_surfels
    before = orc.surfels[:8, :n].copy()
    deleted = orc.end_tasks()
    assert orc.n == n - deleted and deleted >= len(away)
    rows = orc.surfels[:8, :orc.n]
    assert not np.any(rows[0].view(np.uint32) == DELETED)          # compacted: no hole left in [0, surfels_size)
    # which surfels went: recompute the marks from the scratch rows the step leaves behind is not possible after the
    # compaction, so re-run on a copy without compaction effects: a second call is idempotent
    again = orc.end_tasks()
    assert again == 0 and orc.n == n - deleted
    assert np.array_equal(rows.view(np.uint32), orc.surfels[:8, :orc.n].view(np.uint32))
    # the survivors are a subset of the original surfels (positions / normals / descriptors untouched), in the order the
    # reference's compaction produces
    key = lambda a: [tuple(c) for c in a[[0, 1, 2, 3, 6, 7]].view(np.uint32).T]
    orig = {k: i for i, k in enumerate(key(before))}
    src = np.array([orig[k] for k in key(rows)])
    assert len(set(src)) == orc.n
    gone = np.ones(n, bool)
    gone[src] = False
    assert gone[away].all()                                          # unobserved surfels are deleted
    assert gone[front].mean() > 0.9                                  # more free-space violations than observations
    exp = expected_compaction(before, gone)
    assert np.array_equal(exp[[0, 1, 2, 3, 5, 6, 7]].view(np.uint32), rows[[0, 1, 2, 3, 5, 6, 7]].view(np.uint32))
    # radii: the smallest radius^2 measured by an observing keyframe (IEEE half values)
    r2 = rows[4]
    assert np.all(np.isfinite(r2)) and np.all(r2 > 0)
    assert np.array_equal(r2, r2.astype(np.float16).astype(np.float32))


def test_min_observation_count_follows_the_bootstrapping_schedule():
    """direct_ba.h:220-226: 1 below 5 keyframes, 2 below 10, 3 from 10 on."""
    sc, orc, *_ = perturbed_oracle()
    assert sc.cfg.num_keyframes < 5
    n0 = orc.n
    d1 = orc.end_tasks()
    sc2, orc2, *_ = perturbed_oracle()
    orc2.min_observation_counts = (3, 3, 3)
    d3 = orc2.end_tasks()
    assert d3 > d1 > 0 and orc2.n == n0 - d3


def test_bundle_adjustment_runs_the_end_tasks_once_at_the_end():
    sc, orc, *_ = perturbed_oracle()
    n0 = orc.n
    orc.bundle_adjust(True, True, 2, 2)
    assert orc.n == n0 - orc.surfels_deleted and orc.surfels_deleted > 0
    sc2, orc2, *_ = perturbed_oracle()
    orc2.bundle_adjust(True, True, 2, 2, end_tasks=False)
    assert orc2.n == n0


def test_end_tasks_match_reference_cuda_golden():
    """tests/golden/tiny_end_tasks.npz: outputs of the reference's own kernels (tools/make_golden.py::golden_end_tasks)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_end_tasks.npz"))
    sc, orc, *_ = perturbed_oracle()
    assert abs(float(np.sum(sc.surfels[:3, :sc.num_surfels].astype(np.float64))) - float(g["surfel_checksum"])) < 1e-6
    deleted = orc.end_tasks()
    assert deleted == int(g["deleted"]) and orc.n == int(g["surfels_size"])
    assert np.array_equal(orc.surfels[:8, :orc.n].view(np.uint32), g["rows"].view(np.uint32))


def _empty_map_oracle(name="tiny"):
    sc = S.make_scene(S.config_by_name(name))
    orc = O.Oracle(sc)
    orc.poses[:] = sc.poses_true
    orc.n = 0
    return sc, orc


def test_create_surfels_fills_unsupported_cells_once():
    """DirectBA::CreateSurfelsForKeyframe (direct_ba.cc:340-405): one surfel per unsupported sparse cell with a valid pixel; a second
    call finds (almost) every cell supported; the new surfels are associated with their keyframe and carry its measurements."""
    sc, orc = _empty_map_oracle()
    cells = orc.model.cf_w * orc.model.cf_h
    new = orc.create_surfels_for_keyframe(0, filter_new_surfels=False)
    assert 0.8 * cells < new <= cells and orc.n == new
    st = orc.pose_coeffs(0)
    assert st.n_assoc >= 0.98 * new and st.cost_depth < 1e-2 * st.n_assoc      # they lie on the measured depth
    assert orc.create_surfels_for_keyframe(0, filter_new_surfels=False) <= 0.02 * new
    rows = orc.surfels[:8, :orc.n]
    assert np.all(np.isfinite(rows[[0, 1, 2, 4, 6, 7]])) and np.all(rows[4] > 0)
    assert np.array_equal(rows[4], rows[4].astype(np.float16).astype(np.float32))             # radius^2 from the half buffer
    assert np.all(rows[5].view(np.uint32) >> 24 == 0)                                          # uchar4 colour, w = 0
    # the filter (observations in the co-visible keyframes) only removes candidates
    sc2, orc2 = _empty_map_oracle()
    assert 0 < orc2.create_surfels_for_keyframe(0, filter_new_surfels=True) <= new
    # capacity: nothing is created when the buffer would overflow (kernel_create_surfels.cc:163-166)
    sc3, orc3 = _empty_map_oracle()
    orc3.n = orc3.pitch - 10
    assert orc3.create_surfels_for_keyframe(0, filter_new_surfels=False) == 0 and orc3.n == orc3.pitch - 10


def test_merge_removes_duplicates_created_by_other_keyframes():
    sc, orc = _empty_map_oracle("tiny")
    # a buffer with room for two keyframes' worth of surfels
    orc.surfels = np.pad(orc.surfels, ((0, 0), (0, 4096)))
    orc.active = np.zeros(orc.surfels.shape[1], np.uint8)
    orc.pitch = orc.surfels.shape[1]
    n0 = orc.create_surfels_for_keyframe(0, False)
    # the same surface seen again by a keyframe at the same pose yields nothing new; merging keyframe 0 against its own
    # surfels deletes nothing either (one surfel per cell)
    assert orc.merge_surfels_for_keyframe(0) == 0
    # duplicate every surfel with a tiny offset: all the copies that land in a cell with their original are merged
    orc.surfels[:8, n0:2 * n0] = orc.surfels[:8, :n0]
    orc.surfels[0, n0:2 * n0] += 1e-4
    orc.n = 2 * n0
    deleted = orc.merge_surfels_for_keyframe(0)
    assert 0.9 * n0 <= deleted <= n0
    assert orc.compact_surfels() == 2 * n0 - deleted
    assert not np.any(orc.surfels[0, :orc.n].view(np.uint32) == DELETED)


def test_bundle_adjustment_with_surfel_updates_runs_the_lifecycle():
    sc = S.make_scene(S.config_by_name("tiny"))
    orc = O.Oracle(sc)
    orc.surfels = np.pad(orc.surfels, ((0, 0), (0, 8192)))
    orc.active = np.zeros(orc.surfels.shape[1], np.uint8)
    orc.pitch = orc.surfels.shape[1]
    orc.n = sc.num_surfels // 2
    r = orc.bundle_adjust(True, True, 2, 2, do_surfel_updates=True)
    assert r.surfels_created > 0 and r.surfels_merged >= 0
    assert orc.n == r.surfels_size - orc.surfels_deleted and orc.ba_iteration_count == 1
    assert np.all(orc.last_active_in_ba_iteration == 0)
    assert not np.any(orc.surfels[0, :orc.n].view(np.uint32) == DELETED)
