"""GPU parity of the end-of-BA surfel maintenance (PerformBASchemeEndTasks, direct_ba.cc:566-653): the sm_100a path through the
C ABI against the reference's own kernels (oracle/_ref) and the CPU oracle.  Deletion decisions, the compaction
permutation and the radii are integer / exact-value work: the bar is bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from badslam_b200 import scene as S
    from badslam_b200.direct_ba import DirectBA
    from oracle import cpu_oracle, ref_cuda
    assert ref_cuda.available(), "oracle/_ref/libbadslam_ref.so missing (oracle/build_ref.sh)"
    return S, DirectBA, cpu_oracle, ref_cuda


def perturb(sc):
    from badslam_b200.scene import displace_surfels
    return displace_surfels(sc)[0]


@pytest.mark.parametrize("name", ["tiny", "small", "cfg2"])
def test_end_tasks_three_way(mods, name):
    S, DirectBA, O, R = mods
    sc = perturb(S.make_scene(S.config_by_name(name)))
    n = sc.num_surfels
    ba, ref = DirectBA.from_scene(sc), R.RefDirectBA(sc)
    d0, n0 = ba.PerformBASchemeEndTasks()
    d1 = ref.end_tasks()
    assert d0 == d1 > 0 and n0 == ref.surfels_size() == n - d0 == ba.surfels_size()
    a, b = ba.GetSurfelsHost(), ref.surfels()
    assert a.shape == b.shape == (8, n0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))          # same survivors, same slots, same radii
    assert not np.any(a[0].view(np.uint32) == 0x7fffffff)
    if name != "cfg2":
        orc = O.Oracle(sc)
        assert orc.end_tasks() == d0 and orc.n == n0
        assert np.array_equal(orc.surfels[:8, :n0].view(np.uint32), a.view(np.uint32))
    # idempotent
    assert ba.PerformBASchemeEndTasks() == (0, n0)
    assert np.array_equal(ba.GetSurfelsHost().view(np.uint32), a.view(np.uint32))


def test_bundle_adjustment_end_task_schedule(mods):
    """increase_ba_iteration_count = true: end tasks after the iterations; false: before them, once per counter value
    (direct_ba_alternating.cc:313-319,725-735)."""
    S, DirectBA, O, R = mods
    sc = perturb(S.make_scene(S.config_by_name("small")))
    n = sc.num_surfels
    ba, ref = DirectBA.from_scene(sc), R.RefDirectBA(sc)
    ro = ba.BundleAdjustment(None, False, False, False, True, True, 2, 2)
    rr = ref.bundle_adjust(True, True, 2, 2)
    assert ro.surfels_deleted == rr.surfels_deleted > 0 and ro.surfels_size == rr.surfels_size == n - ro.surfels_deleted
    assert ba.ba_iteration_count() == 1
    a, b = ba.GetSurfelsHost(), ref.surfels()
    assert a.shape == b.shape
    # same survivors in the same slots (the two BA iterations before differ by fp32 round-off: a packed normal or a radius
    # decision may flip for a handful of surfels)
    assert (a[3].view(np.uint32) != b[3].view(np.uint32)).mean() < 1e-3 and (a[4] != b[4]).mean() < 1e-3
    assert np.array_equal(a[5].view(np.uint32), b[5].view(np.uint32))
    assert np.abs(a[:3] - b[:3]).mean() < 1e-6 and np.abs(a[:3] - b[:3]).max() < 1e-3   # (displaced surfels that survive are ill-constrained)
    # increase_ba_iteration_count = false: counters 1 != -1 -> end tasks first (nothing left to delete), then not again
    r2 = ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)
    assert r2.surfels_deleted == 0 and ba.last_ba_iteration_count() == 1 and ba.ba_iteration_count() == 1
    ba2 = DirectBA.from_scene(sc)      # a fresh handle (counters 0 != -1) runs the end tasks before its first iteration
    r3 = ba2.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)
    assert r3.surfels_deleted > 0 and ba2.surfels_size() == n - r3.surfels_deleted and ba2.last_ba_iteration_count() == 0


def test_end_tasks_against_golden_fixture(mods):
    """The CUDA path against tests/golden/tiny_end_tasks.npz (outputs of the reference's kernels)."""
    import os
    S, DirectBA, O, R = mods
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_end_tasks.npz"))
    sc = perturb(S.make_scene(S.config_by_name("tiny")))
    ba = DirectBA.from_scene(sc)
    deleted, size = ba.PerformBASchemeEndTasks()
    assert deleted == int(g["deleted"]) and size == int(g["surfels_size"])
    assert np.array_equal(ba.GetSurfelsHost().view(np.uint32), g["rows"].view(np.uint32))
