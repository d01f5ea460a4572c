# ncu full capture of the STT feature kernels (after a plain run of the same command line has exited 0)
set -x
python bench.py --no-extra --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernels_ms_per_step'])
"
python tools/prof_target.py stt 1 > gpurun_out/r02c_plain_stt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_nr_|k_logmel' -c 6 -o gpurun_out/r02c_prof_stt -f python tools/prof_target.py stt 1 > gpurun_out/r02c_ncu.log 2>&1
tail -n 2 gpurun_out/r02c_ncu.log
