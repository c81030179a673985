# First GPU call of the next round: validates what was written after round 1's GPU budget was spent (the fused preprocessing
# kernel), times it, captures it with ncu, and refreshes the headline numbers.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_gpu.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_reference_tests.py tests/test_gpu_parity.py -k "preprocess or reference or progress_function or residual_types or from_buffers" --runxfail -m gpu -q --tb=short -s 2>&1 | grep -v "^E   *+" | cut -c1-400 | tail -60 > gpurun_out/gpu_preprocess_tests.log
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -15 > gpurun_out/gpu_tests.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_preprocess.py -m gpu --runxfail -q -x -k 'ragged or edge' > gpurun_out/sanitizer_preprocess.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitizer_preprocess.log
python tools/make_golden.py --preprocess-only > gpurun_out/golden_preprocess.log 2>&1   # -> gpurun_out/golden/tiny_preprocess.npz, copy to tests/golden/
python tools/preprocess_time.py --size 640x480 > gpurun_out/preprocess_time.log 2>&1
python tools/preprocess_time.py --size 640x480 --flush >> gpurun_out/preprocess_time.log 2>&1
python tools/preprocess_time.py --size 1280x720 >> gpurun_out/preprocess_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'PreprocessFrameKernel' -c 2 -o gpurun_out/r2_preprocess -f python tools/preprocess_time.py --iters 3 > gpurun_out/ncu_preprocess.log 2>&1
python tools/run_dataset.py --make-synthetic /tmp/synth_ds > gpurun_out/run_dataset.log 2>&1 && python tools/run_dataset.py /tmp/synth_ds --keyframe-interval 1 --raw-to-float-depth 0.001 --cell-size 2 --max-depth 6 --max-surfels 1000000 >> gpurun_out/run_dataset.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err
tail -25 gpurun_out/gpu_preprocess_tests.log; tail -4 gpurun_out/sanitizer_preprocess.log; tail -3 gpurun_out/gpu_tests.log; cat gpurun_out/preprocess_time.log; tail -2 gpurun_out/run_dataset.log; tail -c 300 gpurun_out/bench_r2_first.json
