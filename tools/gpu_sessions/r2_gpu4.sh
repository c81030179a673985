set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -40 > gpurun_out/r2_gpu_tests_4.log
python tools/ab_fast.py --steps 5 tools/ab/geo_il1.so tools/ab/geo32.so tools/ab/t320.so tools/ab/t384.so > gpurun_out/r2_ab3.log 2>&1
tail -25 gpurun_out/r2_gpu_tests_4.log; cat gpurun_out/r2_ab3.log
