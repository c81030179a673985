# round 2, GPU call 18: stage counters of the pose kernel as per-lane sums (one warp reduction per chunk instead of two ballots per step)
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -6 > gpurun_out/r2_gpu_tests_18.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-all > gpurun_out/bench_r2_18.json 2> gpurun_out/bench_r2_18.err
tail -3 gpurun_out/r2_gpu_tests_18.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_18.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["ms_per_step","stage_ms","ms_full_ba_10_iterations"]}, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["ms_per_step"])
PY
