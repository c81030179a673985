set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -60 > gpurun_out/r2_gpu_tests_3.log
python tools/ab_fast.py --steps 5 tools/ab/nopacked.so tools/ab/noffma2.so > gpurun_out/r2_ab2.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'PoseAccumulate|ActivationNormals|PositionDescriptor' -c 4 -o gpurun_out/r2_step_a -f python tools/profile_all.py cfg3 > gpurun_out/ncu_step_a.log 2>&1
tail -40 gpurun_out/r2_gpu_tests_3.log; cat gpurun_out/r2_ab2.log; tail -3 gpurun_out/ncu_step_a.log
