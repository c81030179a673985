"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
a 2-rank bundle adjustment (surfel shards + keyframe work list split, NCCL exchange) must reproduce the single-GPU run."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from badslam_b200.direct_ba import DirectBA
    from badslam_b200.scene import config_by_name, make_scene
    sc = make_scene(config_by_name("small"))
    ba = DirectBA.from_scene(sc, device=f"cuda:{rank}", rank=rank, world_size=world)
    ba.SetCollective()
    r = ba.BundleAdjustment(None, False, False, False, True, True, 3, 3)
    poses, act = ba.GetKeyframeStates()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), poses=poses, act=act, surfels=ba.GetSurfelsHost(), active=ba.GetActiveHost(),
             counts=np.array([r.depth_residual_count, r.descriptor_residual_count, r.pose_iterations_total]))
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


def test_two_rank_bundle_adjustment_matches_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from badslam_b200.direct_ba import DirectBA
    from badslam_b200.scene import config_by_name, make_scene, pose_error
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    sc = make_scene(config_by_name("small"))
    ba = DirectBA.from_scene(sc, device="cuda:0")
    r = ba.BundleAdjustment(None, False, False, False, True, True, 3, 3)
    poses, act = ba.GetKeyframeStates()
    surf = ba.GetSurfelsHost()
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # the replicas agree with each other exactly ...
    assert np.array_equal(r0["poses"], r1["poses"]) and np.array_equal(r0["act"], r1["act"])
    assert np.array_equal(r0["surfels"].view(np.uint32), r1["surfels"].view(np.uint32))
    assert np.array_equal(r0["active"], r1["active"])
    # ... and with the single-GPU run up to the summation order of the pose normal equations
    assert tuple(r0["counts"]) == (r.depth_residual_count, r.descriptor_residual_count, r.pose_iterations_total)
    assert np.array_equal(r0["act"], act)
    for k in range(sc.cfg.num_keyframes):
        dt, dr = pose_error(r0["poses"][k], poses[k])
        assert dt < 1e-5 and dr < 1e-5
    assert np.mean(np.abs(r0["surfels"][:3] - surf[:3])) < 1e-6
