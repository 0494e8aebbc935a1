#!/bin/bash
# SASS evidence of the Blackwell-native instructions in the shipped library (no GPU needed):
#   bash tools/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
SO=kornia_b200/_C/libkornia_b200.so
echo "cuobjdump -sass $SO  ($(stat -c %s $SO) bytes, $(date -u +%Y-%m-%dT%H:%MZ)), arch: $(cuobjdump -lelf $SO | head -1)"
TMP=$(mktemp)
cuobjdump -sass $SO > $TMP
echo "kernels (Function :) $(grep -c 'Function :' $TMP)"
for op in UTMALDG UTMAREDG UTMAPF UTMACCTL SYNCS.ARRIVE SYNCS.PHASECHK ELECT FFMA2 LDS.128 STS.128 MUFU.RCP UTCMMA LDTM HMMA; do
  printf "%-16s %s\n" "$op" "$(grep -c "$op" $TMP)"
done
echo
echo "first occurrences:"
for op in UTMALDG UTMAREDG UTMAPF SYNCS.ARRIVE SYNCS.PHASECHK ELECT FFMA2; do grep -m1 "$op" $TMP | sed 's/^ *//' | cut -c1-120; done
echo
echo "per kernel family (UTMALDG / UTMAREDG / FFMA2 counts):"
awk '/Function :/{name=$3} /UTMALDG/{l[name]++} /UTMAREDG/{r[name]++} /FFMA2/{f[name]++} END{for(n in l) print n, l[n]+0, r[n]+0, f[n]+0}' $TMP | sed 's/_ZN5kb200//' | awk '{split($1,a,"I"); fam[a[1]]+=1; L[a[1]]+=$2; R[a[1]]+=$3; F[a[1]]+=$4} END{for(k in fam) printf "%-28s instantiations %3d  UTMALDG %4d  UTMAREDG %3d  FFMA2 %5d\n", k, fam[k], L[k], R[k], F[k]}' | sort
rm -f $TMP
