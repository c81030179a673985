# round 2, GPU call 15 (4 GPUs): 2-rank parity suite after the pose-step / end-task changes, bench at 4 ranks with the replica check
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
nproc > gpurun_out/host_15.log; cat /proc/loadavg >> gpurun_out/host_15.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-500 | tail -25 > gpurun_out/r2_gpu_tests_multi_15.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n4_r2b.json 2> gpurun_out/bench_cfg3_n4_r2b.err
cat gpurun_out/host_15.log; tail -8 gpurun_out/r2_gpu_tests_multi_15.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg3_n4_r2b.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","stage_ms","ms_full_ba_10_iterations","multi_gpu_check","host"]}); print(d.get("e2e")); print(d["roofline"]["avg_launch_ms"], d["roofline"]["kernel_share_of_step"])
PY
tail -3 gpurun_out/bench_cfg3_n4_r2b.err
