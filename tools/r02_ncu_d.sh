set -x
python tools/prof_target.py stt 1 > gpurun_out/r02d_plain_stt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_logmel16|k_nr_istft|k_nr_mask' -c 3 -o gpurun_out/r02d_prof_stt -f python tools/prof_target.py stt 1 > gpurun_out/r02d_ncu.log 2>&1
tail -n 2 gpurun_out/r02d_ncu.log
