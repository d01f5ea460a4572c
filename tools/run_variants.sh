# usage: bash tools/run_variants.sh name1 name2 ... : bench each variants/lib_NAME.so in place of the library (experiments only)
cp open_speech_b200/libosb200.so /tmp/lib_keep.so
for v in "$@"; do cp variants/lib_$v.so open_speech_b200/libosb200.so; echo "== $v"; python tools/bk.py --no-extra --steps 10; done
cp /tmp/lib_keep.so open_speech_b200/libosb200.so
