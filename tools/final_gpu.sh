set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -15 > gpurun_out/gpu_tests.log
python tools/make_golden.py --end-tasks-only > gpurun_out/golden.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r1_final_ref.json 2> gpurun_out/bench_r1_final_ref.err
BADBA_LIB=$PWD/tools/ab/u2.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_u2.json 2> gpurun_out/bench_r1_u2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'PoseAccumulate|ActivationNormals|PositionDescriptor|IntrinsicsAccumulate|PcgAccumulate|ObservationStats' -c 16 -o gpurun_out/r1_final_kernels -f python tools/profile_all.py cfg3 > gpurun_out/ncu_all.log 2>&1
tail -3 gpurun_out/gpu_tests.log; tail -2 gpurun_out/golden.log; tail -c 400 gpurun_out/bench_r1_final.json; tail -2 gpurun_out/ncu_all.log
