# BASELINE.json configs 4 and 5 on ONE B200: parity spot check against the reference's kernels (gated test), size-independent
# checks + timing (tools/run_config.py), the bench line of cfg4 with its reference arm.  Logs -> profiles/bench/.
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
BADBA_BIG_CONFIGS=1 timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k big_config --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-400 | tail -30 > gpurun_out/big_configs_spot_check.log
timeout 900 python tools/run_config.py --workload cfg4 --iterations 3 > gpurun_out/run_cfg4.json 2> gpurun_out/run_cfg4.err
timeout 900 python tools/run_config.py --workload cfg5 --iterations 3 > gpurun_out/run_cfg5.json 2> gpurun_out/run_cfg5.err
timeout 900 python bench.py --workload cfg4 --intrinsics --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-all > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 900 python bench.py --impl reference --workload cfg4 --intrinsics --steps 1 --warmup 1 > gpurun_out/bench_cfg4_ref.json 2> gpurun_out/bench_cfg4_ref.err
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e-all > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 900 python bench.py --impl reference --workload cfg5 --steps 1 --warmup 1 > gpurun_out/bench_cfg5_ref.json 2> gpurun_out/bench_cfg5_ref.err
tail -8 gpurun_out/big_configs_spot_check.log; cat gpurun_out/run_cfg4.json gpurun_out/run_cfg5.json | cut -c1-900; tail -c 700 gpurun_out/bench_cfg4.json; tail -c 400 gpurun_out/bench_cfg4_ref.json; tail -c 700 gpurun_out/bench_cfg5.json; tail -c 400 gpurun_out/bench_cfg5_ref.json; tail -3 gpurun_out/*.err
