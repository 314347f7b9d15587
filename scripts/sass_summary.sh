#!/bin/bash
# Per-kernel SASS evidence of the Blackwell paths in libb200rec.so: tcgen05 MMA (UTCHMMA), TMA loads (UTMALDG),
# TMEM loads (LDTM), legacy mma.sync (HMMA), MUFU.  Usage: scripts/sass_summary.sh > profiles/r02_sass_summary.txt
SO="$(dirname "$0")/../multi-head-recommendation-with-human-priors_b200/libb200rec.so"
echo "# cuobjdump -sass $(basename "$SO") ($(date -u +%Y-%m-%dT%H:%MZ)), sm_100a; counts of instruction mnemonics per kernel"
cuobjdump -sass "$SO" | awk '
  /Function : / { fn=$3; next }
  /UTCHMMA/ { mma[fn]++; if ($0 ~ /2CTA/) mma2[fn]++ }
  /UTMALDG/ { tma[fn]++ }
  /UTMASTG/ { tmas[fn]++ }
  /LDTM/ { ldtm[fn]++ }
  /STTM/ { sttm[fn]++ }
  /HMMA/ && !/UTCHMMA/ { hmma[fn]++ }
  /MUFU/ { mufu[fn]++ }
  /SYNCS/ { syncs[fn]++ }
  { total[fn]++ }
  END {
    printf "%-8s %-6s %-8s %-8s %-6s %-6s %-6s %-6s %-8s %s\n", "UTCHMMA", "2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "MUFU", "instrs", "kernel"
    for (f in total) if (mma[f] + tma[f] + ldtm[f] + hmma[f] > 0)
      printf "%-8d %-6d %-8d %-8d %-6d %-6d %-6d %-6d %-8d %s\n", mma[f], mma2[f], tma[f], tmas[f], ldtm[f], sttm[f], hmma[f], mufu[f], total[f], f
  }' | (read -r hdr; echo "$hdr"; sort -k10)
