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
