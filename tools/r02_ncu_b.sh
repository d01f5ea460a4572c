# round-2 ncu evidence, second pass (after the recurrence change): the composed chain's launch list and the VAD kernels' full capture
set -x
python tools/prof_target.py full 2 > gpurun_out/r02_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_stt_full.csv python tools/prof_target.py full 2 > gpurun_out/r02_ncu_b.log 2>&1
python tools/prof_target.py vad 1 > gpurun_out/r02_plain_vad.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_vad_front_fused|k_vad_recur' -c 2 -o gpurun_out/r02_prof_vad -f python tools/prof_target.py vad 1 > gpurun_out/r02_ncu_d.log 2>&1
for f in gpurun_out/r02_ncu_[bd].log; do tail -n 2 $f; done
