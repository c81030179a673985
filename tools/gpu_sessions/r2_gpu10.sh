# round 2, GPU call 10 (2 GPUs): multi-rank parity (BA, intrinsics, surfel updates, PCG, sharded end tasks), odometry tests, 2-GPU bench line
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_odometry.py -m gpu -q -s --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-600 | tail -120 > gpurun_out/r2_gpu_tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_cfg3_n2_r2.json 2> gpurun_out/bench_cfg3_n2_r2.err
head -110 gpurun_out/r2_gpu_tests_multi.log; tail -c 1800 gpurun_out/bench_cfg3_n2_r2.json; tail -5 gpurun_out/bench_cfg3_n2_r2.err
