"""NVLink traffic of the fused geometry exchange (peer stores from the geometry kernels), measured with the NVML throughput
counters around a run of BA iterations on 2 GPUs, next to the byte count the store pattern implies.

    python tools/nvlink_bytes.py [--workload cfg2] [--steps 200]        (spawns 2 ranks; needs 2 GPUs)
"""
import argparse
import json
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nvlink_kib(index):
    """(tx KiB, rx KiB) summed over the links of GPU `index`, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        tx_id, rx_id = pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX
        try:
            vals = pynvml.nvmlDeviceGetFieldValues(h, [(tx_id, 0xFFFFFFFF), (rx_id, 0xFFFFFFFF)])   # scope: all links
            if all(v.nvmlReturn == 0 for v in vals):
                return int(vals[0].value.ullVal), int(vals[1].value.ullVal)
        except Exception:
            pass
        tx = rx = 0
        seen = False
        for link in range(18):
            try:
                vals = pynvml.nvmlDeviceGetFieldValues(h, [(tx_id, link), (rx_id, link)])
            except Exception:
                break
            if vals[0].nvmlReturn == 0 and vals[1].nvmlReturn == 0:
                tx += int(vals[0].value.ullVal)
                rx += int(vals[1].value.ullVal)
                seen = True
        return (tx, rx) if seen else None
    except Exception as e:   # noqa: BLE001
        print("nvml unavailable:", type(e).__name__, e, flush=True)
        return None


def worker(rank, world, port, workload, steps, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from badslam_b200.direct_ba import DirectBA
    from badslam_b200.scene import config_by_name, make_scene
    sc = make_scene(config_by_name(workload))
    K = sc.cfg.num_keyframes
    ba = DirectBA.from_scene(sc, device=f"cuda:{rank}", rank=rank, world_size=world)
    ba.SetCollective()
    peers = ba.EnablePeerExchange()
    surf = ba.surfels()
    backup = surf[:8].clone()
    poses0, act0 = sc.poses_init.copy(), np.zeros(K, np.int32)
    ba.SetLastBAIterationCount(ba.ba_iteration_count())

    def step():
        surf[:8].copy_(backup, non_blocking=True)
        ba.MarkReplicaRewritten()
        ba.SetKeyframeStates(poses0, act0)
        return ba.BundleAdjustment(None, False, False, False, True, True, 1, 1, increase_ba_iteration_count=False)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier(device_ids=[rank])
    before = nvlink_kib(rank)
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dist.barrier(device_ids=[rank])
    after = nvlink_kib(rank)
    active = int(ba.GetActiveHost().sum())
    if rank == 0:
        n = sc.num_surfels
        # what the kernels store into the ONE peer replica per iteration: this rank's surfels only (half of them): packed
        # normal (4 B) + active flag (1 B) from the activation / normals pass, x y z d1 d2 (20 B) from the position / descriptor pass
        expected = (n / world) * (world - 1) * 25.0
        out = {"workload": workload, "ranks": world, "peers_mapped": peers, "steps": steps, "surfels": n, "active_surfels": active,
               "expected_store_bytes_per_iteration_per_rank": expected}
        if before and after:
            out["nvlink_tx_bytes_per_iteration"] = (after[0] - before[0]) * 1024.0 / steps
            out["nvlink_rx_bytes_per_iteration"] = (after[1] - before[1]) * 1024.0 / steps
        else:
            out["nvlink_counters"] = "unavailable"
        with open(out_path, "w") as f:
            json.dump(out, f)
        print(json.dumps(out), flush=True)
    dist.barrier(device_ids=[rank])
    dist.destroy_process_group()


def main():
    import torch
    import torch.multiprocessing as mp
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    if torch.cuda.device_count() < 2:
        print("needs 2 GPUs")
        return 1
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    mp.spawn(worker, args=(2, port, a.workload, a.steps, os.path.join(ROOT, "gpurun_out", "nvlink_bytes.json")), nprocs=2, join=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
