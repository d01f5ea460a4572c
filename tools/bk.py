"""python tools/bk.py [bench args]: run bench.py and print ms/step + per-kernel ms (one line)."""
import json, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", *sys.argv[1:]], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(round(d["ms_per_step"], 3), d.get("roofline", {}).get("kernels_ms_per_step"), flush=True)
except Exception as e:
    print("bench failed", e, out.stdout[-500:], out.stderr[-1500:])
