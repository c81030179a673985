# round 2, GPU call 8: image-pair odometry -- three-way tests, golden fixture, timing, launch list + ncu of the persistent kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_odometry.py -m gpu -q -s --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-500 | tail -60 > gpurun_out/r2_gpu_tests_odometry.log
timeout 300 python tools/make_golden.py --odometry-only > gpurun_out/golden_odometry.log 2>&1
timeout 300 python tools/odometry_time.py > gpurun_out/odometry_time.log 2>&1
timeout 300 python tools/odometry_time.py --size 320x240 --scales 4 >> gpurun_out/odometry_time.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'OdomTrack|Level0Kernel|DownsampleKernel|BrightnessKernel' -c 8 -o gpurun_out/r2_odometry -f python tools/odometry_time.py --reps 1 > gpurun_out/ncu_odometry.log 2>&1
tail -40 gpurun_out/r2_gpu_tests_odometry.log; tail -3 gpurun_out/golden_odometry.log; cat gpurun_out/odometry_time.log; tail -2 gpurun_out/ncu_odometry.log
