#!/bin/bash
# Opcode histogram of the Blackwell-specific instructions per kernel of the built library (what proves tcgen05 / TMA / TMEM:
# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA loads / stores; /opt/skills/guides/B200_PROFILING.md).
#   tools/sass_histogram.sh [lib] > profiles/rNN_sass_opcode_histogram.txt
LIB=${1:-streamyolo_b200/lib/libstreamyolo_sm100.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ { fn=$3; next }
/\/\*[0-9a-f]+\*\// {
  for (i = 1; i <= NF; ++i) if ($i ~ /^(@!?U?P[0-9T]+)$/) continue; else if ($i ~ /^[A-Z][A-Za-z0-9_.]+;?$/) { op=$i; break }
  gsub(/;/, "", op);
  if (op ~ /^(UTC|UTMA|LDTM|STTM|HMMA|LDSM|UBLKCP|SYNCS|ELECT|UTCBAR)/) c[fn "\t" op]++
}
END { for (k in c) print c[k] "\t" k }' | sort -t$'\t' -k2,2 -k1,1nr | c++filt 2>/dev/null | awk -F'\t' '{printf "%6d  %-28s %s\n", $1, $3, substr($2, 1, 110)}'
