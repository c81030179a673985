# round 2, GPU call 9: odometry tests + golden fixture, PCG regression (fixed-order sums), locality A/B of the pixel gathers
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_odometry.py -m gpu -q -s --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-600 | tail -150 > gpurun_out/r2_gpu_tests_odometry.log
timeout 300 python tools/make_golden.py --odometry-only > gpurun_out/golden_odometry.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_tests.py -m gpu -q --tb=short -k "pcg or PCG" 2>&1 | grep -v "^E   *+" | cut -c1-400 | tail -40 > gpurun_out/r2_gpu_tests_pcg.log
timeout 900 python tools/ab_locality.py --workload cfg3 --steps 5 > gpurun_out/ab_locality.log 2>&1
head -100 gpurun_out/r2_gpu_tests_odometry.log; tail -4 gpurun_out/golden_odometry.log; tail -8 gpurun_out/r2_gpu_tests_pcg.log; cat gpurun_out/ab_locality.log
