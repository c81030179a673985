# round 2, GPU call 13: A/B of the small-shard geometry changes (records staged once per CTA, keyframes per work item) on the workload
# one rank of an 8-GPU job sees and on cfg3, A/B of the per-item chunk size (Gauss-Newton tail) on cfg3, parity of the changed kernels
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --tb=short -x 2>&1 | grep -v "^E   *+" | cut -c1-400 | tail -15 > gpurun_out/r2_gpu_tests_13.log
timeout 900 python tools/ab_fast.py --workload cfg3_rank8 --steps 10 tools/ab/nostageall.so env:BADBA_GEO_GROUP=32 env:BADBA_GEO_GROUP=64 env:BADBA_GEO_GROUP=200 > gpurun_out/r2_ab_rank8.log 2>&1
timeout 900 python tools/ab_fast.py --workload cfg3 --steps 5 tools/ab/launchchunks.so tools/ab/nostageall.so env:BADBA_GEO_GROUP=32 env:BADBA_GEO_GROUP=64 > gpurun_out/r2_ab_cfg3.log 2>&1
tail -6 gpurun_out/r2_gpu_tests_13.log; cat gpurun_out/r2_ab_rank8.log gpurun_out/r2_ab_cfg3.log
