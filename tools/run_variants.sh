# usage: bash tools/run_variants.sh "bench args" name1 name2 ... : bench each variants/lib_NAME.so in place of the library (experiments only)
args=$1; shift
cp open_speech_b200/libosb200.so /tmp/lib_keep.so
for v in "$@"; do cp variants/lib_$v.so open_speech_b200/libosb200.so; echo "== $v"; python tools/bk.py $args; done
cp /tmp/lib_keep.so open_speech_b200/libosb200.so
