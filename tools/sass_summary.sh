#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses (B200_PROFILING.md): run here, no GPU needed.
# usage: tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=open_speech_b200/libosb200.so
echo "# cuobjdump -sass $LIB (sm_100a).  tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, cp.async.bulk (TMA unit) = UBLKCP,"
echo "# mbarrier = SYNCS, packed FP32 (F32x2 column) = FFMA2 + FADD2 + FMUL2, FP64 = DFMA/DADD/DMUL, warp-level MMA = HMMA (only k_vad_recur_tc, the eight-stream LSTM recurrence: DESIGN 4.5)"
printf "%-58s %8s %6s %7s %7s %6s %6s %7s %6s %5s\n" kernel UTCHMMA LDTM UTCBAR UBLKCP SYNCS F32x2 FFMA F64 HMMA
cuobjdump -sass $LIB | awk '
/Function :/ { if (name != "") out(); name=$3; u=l=b=k=s=f2=f=d=h=0; next }
/UTCHMMA/ {u++} /LDTM/ {l++} /UTCBAR/ {b++} /UBLKCP/ {k++} /SYNCS/ {s++} /FFMA2|FADD2|FMUL2/ {f2++} / FFMA / {f++} /DFMA|DADD|DMUL/ {d++} / HMMA/ {h++}
function out() { print name, u, l, b, k, s, f2, f, d, h }
END { out() }' | while read -r n u l b k s f2 f d h; do
    dn=$(echo "$n" | c++filt | sed 's/(.*//; s/^void //; s/osb:://g' | cut -c1-58)
    printf "%-58s %8d %6d %7d %7d %6d %6d %7d %6d %5d\n" "$dn" $u $l $b $k $s $f2 $f $d $h
done | sort
