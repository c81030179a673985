# round 2, GPU call 6 (state check after the container was re-created): GPU suite, bench both arms, launch list, ncu of the step
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -40 ) > gpurun_out/r2_gpu_tests_6.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_6.json 2> gpurun_out/bench_r2_6.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2_6_ref.json 2> gpurun_out/bench_r2_6_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_6.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-all > gpurun_out/ncu_bench_6.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'PoseAccumulate|ActivationNormals|PositionDescriptor|ObservationStats' -c 10 -o gpurun_out/r2_step_6 -f python tools/profile_all.py cfg3 > gpurun_out/ncu_step_6.log 2>&1
tail -12 gpurun_out/r2_gpu_tests_6.log; tail -c 1500 gpurun_out/bench_r2_6.json; tail -c 600 gpurun_out/bench_r2_6_ref.json; tail -3 gpurun_out/ncu_step_6.log
