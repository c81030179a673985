"""A/B: runs bench.py once per library variant (BADBA_LIB) and prints the step time + stage split of each."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = [("in-tree", None)] + [(os.path.basename(p), os.path.join(root, p)) for p in sys.argv[1:]]
for name, path in variants:
    env = dict(os.environ)
    if path:
        env["BADBA_LIB"] = path
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline"],
                         env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(name, "FAILED", out.stderr[-500:])
        continue
    j = json.loads(line[-1])
    print(f"{name}: {j['ms_per_step']:.2f} ms/step value {j['value']:.3e} roofline {j['roofline']['frac']:.3f} config {json.dumps(j['config'])[:400]}", flush=True)
