# round-2 ncu evidence (one gpurun call; every ncu command follows a plain run of the same command line that exited 0)
set -x
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_plain_default.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_ncu_a.log 2>&1
python tools/prof_target.py full 2 > gpurun_out/r02_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_stt_full.csv python tools/prof_target.py full 2 > gpurun_out/r02_ncu_b.log 2>&1
python tools/prof_target.py stt 1 > gpurun_out/r02_plain_stt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_nr_|k_logmel|k_mel_gain' -c 7 -o gpurun_out/r02_prof_stt -f python tools/prof_target.py stt 1 > gpurun_out/r02_ncu_c.log 2>&1
python tools/prof_target.py full 1 > gpurun_out/r02_plain_full1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_resample_linear_tiled' -c 1 -o gpurun_out/r02_prof_resample -f python tools/prof_target.py full 1 > gpurun_out/r02_ncu_f.log 2>&1
python tools/prof_target.py vad 1 > gpurun_out/r02_plain_vad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_vad_front_fused|k_vad_recur' -c 2 -o gpurun_out/r02_prof_vad -f python tools/prof_target.py vad 1 > gpurun_out/r02_ncu_d.log 2>&1
python tools/prof_target.py tts 1 > gpurun_out/r02_plain_tts.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_fx_reverb_eq' -c 1 -o gpurun_out/r02_prof_tts -f python tools/prof_target.py tts 1 > gpurun_out/r02_ncu_e.log 2>&1
for f in gpurun_out/r02_ncu_?.log; do tail -n 2 $f; done
