set -x
python tools/prof_target.py tts 1 > gpurun_out/r02_plain_tts.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_fx_reverb_eq' -c 1 -o gpurun_out/r02_prof_tts -f python tools/prof_target.py tts 1 > gpurun_out/r02_ncu_e.log 2>&1
tail -n 2 gpurun_out/r02_ncu_e.log
