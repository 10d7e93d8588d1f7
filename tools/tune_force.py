"""Force-kernel tuning sweep on the LJ-1M bench workload (variants defined in engine.cu: EMDEE_TUNE_CASE)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in range(12):
    env = dict(os.environ, EMDEE_FORCE_TUNE=str(v))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "10", "--no-cpu-baseline"],
                       capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(v, "force_ms=%.4f" % d["timing"]["force_kernel_ms"], "step_ms=%.4f" % d["ms_per_step"], flush=True)
    except Exception as e:
        print(v, "FAILED", r.stderr[-300:], flush=True)
