# refresh of the profiles/ evidence for the default bench line (run after the plain bench has exited 0)
python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_nr_|k_logmel' -c 12 -o gpurun_out/prof_r01h -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python bench.py --workload tts --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_tts.json
