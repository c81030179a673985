# round 2, GPU call 2: full GPU suite incl. the full-size parity tests, A/B of pose-kernel variants, preprocessing ncu
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s -x 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -60 > gpurun_out/r2_gpu_tests_2.log
python tools/ab_fast.py --steps 5 tools/ab/r1.so tools/ab/noffma2.so tools/ab/chunk9.so tools/ab/chunk10.so > gpurun_out/r2_ab1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_preprocess_launches.csv python tools/preprocess_time.py --iters 3 > gpurun_out/ncu_pre_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:PreprocessFrameKernel -c 1 -o gpurun_out/r2_preprocess -f python tools/preprocess_time.py --iters 3 > gpurun_out/ncu_preprocess.log 2>&1
tail -30 gpurun_out/r2_gpu_tests_2.log; cat gpurun_out/r2_ab1.log; grep -v "^==" gpurun_out/r2_preprocess_launches.csv | cut -d, -f5,12- | tail -30
