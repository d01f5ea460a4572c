# launch lists refreshed after the tensor-pipe recurrence (one gpurun call; every ncu command follows a plain run of the same command line that exited 0)
set -x
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_ncu_a.log 2>&1
python tools/prof_target.py full 2 > gpurun_out/r02_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_stt_full.csv python tools/prof_target.py full 2 > gpurun_out/r02_ncu_b.log 2>&1
for f in gpurun_out/r02_ncu_[ab].log; do tail -n 2 $f; done
