"""Print the per-kernel CUDA-event times of one bench configuration (tuning helper)."""
import json, subprocess, sys, os
env = dict(os.environ)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    env[k] = v
out = subprocess.run([sys.executable, "bench.py", "--steps", "50", "--warmup", "5", "--no-cpu-baseline"], capture_output=True, text=True, env=env).stdout
l = [x for x in out.splitlines() if x.startswith("{")]
d = json.loads(l[-1])
print(" ".join(sys.argv[1:]) or "default", "| ms/step %.3f |" % d["ms_per_step"], {k: round(v * 1e3) for k, v in d["roofline"]["kernel_ms"].items()})
