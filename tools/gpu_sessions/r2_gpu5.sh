set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -40 > gpurun_out/r2_gpu_tests_5.log
python tools/ab_fast.py --steps 5 env:BADBA_POSE_NO_PRECOMPUTE=1 tools/ab/bar.so tools/ab/geopacked.so > gpurun_out/r2_ab4.log 2>&1
tail -25 gpurun_out/r2_gpu_tests_5.log; cat gpurun_out/r2_ab4.log
