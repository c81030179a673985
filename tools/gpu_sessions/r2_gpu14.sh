# round 2, GPU call 14: parity suite of the BA kernels after the geometry / pose-tail changes, A/B against the previous behaviour
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_lifecycle.py -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-400 | tail -25 > gpurun_out/r2_gpu_tests_14.log
timeout 900 python tools/ab_fast.py --workload cfg3_rank8 --steps 10 tools/ab/nostageall.so > gpurun_out/r2_ab_rank8_b.log 2>&1
timeout 900 python tools/ab_fast.py --workload cfg3 --steps 8 tools/ab/launchchunks.so tools/ab/nostageall.so > gpurun_out/r2_ab_cfg3_b.log 2>&1
tail -12 gpurun_out/r2_gpu_tests_14.log; cat gpurun_out/r2_ab_rank8_b.log gpurun_out/r2_ab_cfg3_b.log | cut -c1-500
