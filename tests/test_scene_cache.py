"""The seeded scene generator: determinism and the optional pickle cache (BADBA_SCENE_CACHE) that lets the processes of one GPU
session -- tests, tools, the ranks of a multi-GPU bench -- generate a configuration once."""
import numpy as np

from badslam_b200 import scene as S


def test_scene_is_deterministic_and_cache_round_trips(tmp_path, monkeypatch):
    cfg = S.config_by_name("tiny")
    a = S.make_scene(cfg)
    monkeypatch.setenv("BADBA_SCENE_CACHE", str(tmp_path))
    b = S.make_scene(cfg)                       # generated, then written
    path = S.scene_cache_path(cfg, str(tmp_path))
    assert path.endswith(".pkl") and (tmp_path / path.split("/")[-1]).exists()
    c = S.make_scene(cfg)                       # loaded
    for x in (b, c):
        assert x.num_surfels == a.num_surfels
        for name in ("depth", "normals", "radius", "color", "surfels", "poses_init", "poses_true", "cfactor"):
            assert np.array_equal(getattr(x, name), getattr(a, name)), name
    # a different configuration gets its own file
    other = S.SceneConfig(width=96, height=64, num_keyframes=2, num_surfels=500, cell=2, seed=5, name="tiny")
    assert S.scene_cache_path(other, str(tmp_path)) != path


def test_render_frame_matches_keyframe_rendering():
    """render_frame (the input of frame tracking / odometry tests) at a keyframe's own pose reproduces that keyframe's images."""
    sc = S.make_scene(S.config_by_name("tiny"))
    d, n, r, c = S.render_frame(sc, sc.poses_true[1])
    assert np.array_equal(d, sc.depth[1]) and np.array_equal(n, sc.normals[1]) and np.array_equal(r, sc.radius[1])
    assert np.array_equal(c, sc.color[1])


def _rank(rank, cache_dir, out):
    import os
    import sys
    import time
    os.environ["BADBA_SCENE_CACHE"] = cache_dir
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    if rank == 0:
        time.sleep(1.0)          # the other rank is already waiting for the pickle
    t0 = time.time()
    sc = bench.load_scene("tiny", rank, 2, wait_seconds=60)
    out.put((rank, sc.num_surfels, float(sc.surfels[:3, :sc.num_surfels].astype("float64").sum()), time.time() - t0))


def test_bench_ranks_share_one_generated_scene(tmp_path):
    """bench.py with one process per GPU: rank 0 generates, the others wait for its pickle and load it."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank, args=(r, str(tmp_path), q)) for r in (1, 0)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [g[0] for g in got] == [0, 1] and got[0][1:3] == got[1][1:3]
    assert got[1][3] >= 0.5            # rank 1 did wait for rank 0
    assert len(list(tmp_path.glob("tiny_*.pkl"))) == 1
