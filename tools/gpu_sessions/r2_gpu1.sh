# round 2, first GPU call: honest GPU suite + preprocessing diagnostics + golden + timing + baseline bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^E   *+" | cut -c1-400 | tail -120 > gpurun_out/r2_gpu_tests_1.log
python tools/preprocess_diag.py > gpurun_out/r2_preprocess_diag.log 2>&1
python tools/make_golden.py --preprocess-only > gpurun_out/golden_preprocess.log 2>&1
python tools/preprocess_time.py --size 640x480 > gpurun_out/preprocess_time.log 2>&1
python tools/preprocess_time.py --size 640x480 --flush >> gpurun_out/preprocess_time.log 2>&1
python tools/preprocess_time.py --size 1280x720 >> gpurun_out/preprocess_time.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err
tail -40 gpurun_out/r2_gpu_tests_1.log; cat gpurun_out/r2_preprocess_diag.log; cat gpurun_out/preprocess_time.log; tail -c 600 gpurun_out/bench_r2_first.json
