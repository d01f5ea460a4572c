set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
python bench.py --workload tts --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_tts.json
python bench.py --workload c1 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_c1.json
python bench.py --workload vad --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_vad.json
python bench.py --workload realtime 2>/dev/null | tail -1 > gpurun_out/bench_realtime.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for f in reference default tts c1 vad realtime; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_$f.json"))
print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
PY
done
