# round 2, final state on one B200: the whole GPU suite, smoke, both bench arms, launch list, ncu capture of the step
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -30 ) > gpurun_out/r2_final_gpu_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e-all > gpurun_out/ncu_bench_final.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'PoseAccumulate|ActivationNormals|PositionDescriptor|ObservationStats' -c 10 -o gpurun_out/r2_final_step -f python tools/profile_all.py cfg3 > gpurun_out/ncu_step_final.log 2>&1
tail -8 gpurun_out/r2_final_gpu_tests.log; cat gpurun_out/r2_final_smoke.log | tail -2; tail -c 1600 gpurun_out/bench_r2_final.json; tail -c 500 gpurun_out/bench_r2_final_ref.json; tail -2 gpurun_out/ncu_step_final.log
