#!/usr/bin/env bash
# SASS evidence of the Blackwell-native paths in the shipped .so (runs without a GPU): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM,
# TMA tensor loads / stores -> UTMALDG / UTMASTG, bulk copies -> UBLKCP; legacy mma.sync would show as a bare HMMA.
cd "$(dirname "$0")/.."
SO=ns2vc_b200/_C/libns2vc_b200.so
T=$(mktemp); cuobjdump -sass $SO > $T
{
  echo "# ${1:-r02} SASS evidence: cuobjdump -sass $SO (commit $(git rev-parse --short HEAD))"
  for m in UTCHMMA UTCQMMA LDTM STTM UTMALDG UTMASTG UBLKCP UTCBAR; do printf "%-10s %s\n" $m "$(grep -c "$m" $T)"; done
  printf "%-10s %s   (legacy mma.sync; UTCHMMA lines excluded)\n" HMMA "$(grep "HMMA" $T | grep -vc UTCHMMA)"
  echo "# per kernel (UTCHMMA / LDTM / UTMALDG / UTMASTG):"
  awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTMALDG/ {c[name]++} /UTMASTG/ {d[name]++} END {for (n in a) printf "%s %d %d %d %d\n", n, a[n], b[n], c[n], d[n]}' $T | c++filt | sed 's/ns2vc:://g; s/(anonymous namespace):://g' | sort
} > profiles/${1:-r02}_sass_counts.txt
rm -f $T
cat profiles/${1:-r02}_sass_counts.txt
