# round 2, GPU call 17: HEAD check -- smoke (with the odometry leg), the BA parity tests, a bench line (pose-step queue depth 3)
set -x
export BADBA_SCENE_CACHE=/tmp/badba_scenes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_head_smoke.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_odometry.py -m gpu -q --tb=short 2>&1 | grep -v "^E   *+" | cut -c1-300 | tail -8 > gpurun_out/r2_head_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-all > gpurun_out/bench_r2_head.json 2> gpurun_out/bench_r2_head.err
tail -3 gpurun_out/r2_head_smoke.log; tail -4 gpurun_out/r2_head_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_head.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["ms_per_step","stage_ms","ms_full_ba_10_iterations","gpu_launches","host"]}, d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["e2e"]["ms_per_step"])
PY
