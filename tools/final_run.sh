# Final evidence of a round, one gpurun call: full GPU test suite, reference arm, every bench line, smoke().  Outputs -> gpurun_out/r02f_*
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02f_bench_reference.json
python bench.py --impl reference --workload stt_full --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02f_bench_reference_stt_full.json
python bench.py --steps 10 --warmup 3 2>gpurun_out/r02f_bench_default.err | tail -1 > gpurun_out/r02f_bench_default.json
python bench.py --workload stt_full --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02f_bench_stt_full.json
python bench.py --workload c1 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02f_bench_c1.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for f in reference reference_stt_full default stt_full c1; do python - <<PY
import json
d=json.load(open("gpurun_out/r02f_bench_$f.json"))
print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("kernel_frac"))
if "configs" in d: print({k: (round(v.get("value", 0)), v.get("ms_per_step")) for k, v in d["configs"].items()})
PY
done
python bench.py --workload tts --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02f_bench_tts.json
python bench.py --workload vad --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02f_bench_vad.json
python bench.py --workload realtime 2>/dev/null | tail -1 > gpurun_out/r02f_bench_realtime.json
for f in tts vad realtime; do python - <<PY
import json
d=json.load(open("gpurun_out/r02f_bench_$f.json"))
print("$f", d.get("value"), d.get("ms_per_step"))
PY
done
