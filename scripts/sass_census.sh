#!/bin/bash
# SASS census of the built library: per kernel, how many tcgen05 / TMA / TMEM / mbarrier instructions the binary holds
# (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
#  UTCATOMSWS = TMEM allocator, SYNCS = mbarrier ops, REDUX = elect / warp reduce, ACQBULK = griddepcontrol.wait,
#  HMMA = legacy mma.sync: must be absent; the .2CTA forms are the cta_group::2 instructions of the CTA-pair kernel).
# usage: scripts/sass_census.sh [path/to/libb200lic.so] > profiles/r2_sass_census.txt
LIB=${1:-rdo_ptq_b200/lib/libb200lic.so}
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn = $3 }
  { for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTMALDG|UTMASTG|UTCBAR|LDTM|UTCATOMSWS|SYNCS|REDUX|HMMA|UBLKCP|ACQBULK)/) { split($i, b, "."); op = b[1]; if ($i ~ /2CTA/) op = op ".2CTA"; c[fn "\t" op]++ } }
  END { for (k in c) print k "\t" c[k] }' | while IFS=$'\t' read -r fn op n; do
    name=$(echo "$fn" | c++filt | sed -E 's/^void //; s/\(.*$//; s/^b200lic:://')
    printf "%s\t%s\t%s\n" "$name" "$op" "$n"
  done | sort -t$'\t' -k1,1 -k3,3nr |
  awk -F'\t' '{ if ($1 != last) { if (last != "") print line; line = $1 ":"; last = $1 } line = line " " $2 "=" $3 } END { print line }'
