"""Host-side logic of the multi-GPU path on CPU: the surfel / keyframe partition exposed by the C ABI and the two
exchange patterns (in-place all-gather of shards, sum all-reduce over disjoint slots), exercised with world_size 2
over the gloo backend."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(lib, n, rank, world):
    b, e = C.c_uint32(), C.c_uint32()
    lib.bba_shard_surfel_range(n, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 1000, 200_000, 3_000_000])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_surfel_shards_partition_the_range(n, world):
    from badslam_b200 import _lib
    lib = _lib.load()
    covered = 0
    prev_end = 0
    for r in range(world):
        b, e = shard(lib, n, r, world)
        assert b == prev_end or (b == n and e == n)       # contiguous, in rank order
        assert b % 256 == 0 or b == n                      # tile aligned (the geometry kernels work on 256-surfel tiles)
        assert e <= n
        covered += e - b
        prev_end = max(prev_end, e)
    assert covered == n


def test_keyframe_work_list_is_dealt_round_robin():
    from badslam_b200 import _lib
    lib = _lib.load()
    for world in (1, 2, 8):
        owners = [lib.bba_shard_keyframe_owner(i, world) for i in range(37)]
        assert owners == [i % world if world > 1 else 0 for i in range(37)]
        counts = np.bincount(owners, minlength=world)
        assert counts.max() - counts.min() <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from badslam_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)                      # identical replicated state on every rank
    rows = rng.normal(size=(7, n)).astype(np.float32)
    new_rows = rng.normal(size=(7, n)).astype(np.float32)   # what a full (single-rank) geometry step would produce
    b, e = shard(lib, n, rank, world)
    local = rows.copy()
    local[:, b:e] = new_rows[:, b:e]                    # this rank only updates its shard
    # exchange buffer: world slices of [7][shard_len] (kernels.cu PackShardKernel / UnpackShardsKernel)
    shard_len = ((n + 255) // 256 + world - 1) // world * 256
    buf = torch.zeros(world * 7 * shard_len)
    sl = buf.view(world, 7, shard_len)
    sl[rank, :, :e - b] = torch.from_numpy(local[:, b:e])
    dist.all_gather_into_tensor(buf, buf.view(world, -1)[rank].clone())
    merged = local.copy()
    for r in range(world):
        rb, re = shard(lib, n, r, world)
        if r != rank:
            merged[:, rb:re] = buf.view(world, 7, shard_len)[r, :, :re - rb].numpy()
    assert np.array_equal(merged, new_rows)             # every replica equals the single-rank result, bit for bit

    # pose slots: each keyframe of the work list is owned by exactly one rank; sum over disjoint slots == gather
    work = list(range(0, K, 2)) + [1]                    # some non-inactive keyframes
    truth = rng.normal(size=(K, 17)).astype(np.float32)
    slots = torch.zeros(K, 17)
    for i, kf in enumerate(work):
        if lib.bba_shard_keyframe_owner(i, world) == rank:
            slots[kf] = torch.from_numpy(truth[kf])
    dist.all_reduce(slots)
    for kf in work:
        assert np.array_equal(slots[kf].numpy(), truth[kf])
    untouched = [k for k in range(K) if k not in work]
    assert not slots[untouched].any()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_exchange_patterns_world_size_2_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 1000, 9, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
