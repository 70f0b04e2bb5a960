#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): tcgen05.mma = UTC*MMA,
# tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG, legacy mma.sync = HMMA.  Runs on the build box (no GPU needed).
#   tools/sass_summary.sh > profiles/r02_sass_summary.txt
so=${1:-moleculediffusiontransformer_b200/libmdt_b200.so}
echo "# cuobjdump -sass $so  ($(date -u +%F), $(nvcc --version | tail -1))"
cuobjdump -sass "$so" | awk '
  /Function :/ { fn=$3 }
  /UTC[A-Z]*MMA/ { u[fn]++ } /LDTM/ { l[fn]++ } /STTM/ { s[fn]++ } /UTMALDG/ { t[fn]++ } /UTMASTG/ { ts[fn]++ }
  / HMMA\./ { h[fn]++ } /SYNCS/ { y[fn]++ } /UTCBAR/ { b[fn]++ } /FENCE.VIEW.ASYNC/ { f[fn]++ }
  END { printf "%-8s %-6s %-6s %-8s %-8s %-6s %-7s %-6s  %s\n", "UTCxMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "UTCBAR", "FENCE", "kernel";
        for (k in y) if (u[k] + l[k] + t[k] + h[k] > 0) printf "%-8d %-6d %-6d %-8d %-8d %-6d %-7d %-6d  %s\n", u[k], l[k], s[k], t[k], ts[k], h[k], b[k], f[k], k }' | (read -r hdr; echo "$hdr"; sort -k9) | while read -r line; do
    name=$(echo "$line" | awk '{print $9}'); dem=$(echo "$name" | c++filt 2>/dev/null | cut -c1-110); echo "$line" | awk -v d="$dem" '{ $9=d; print }' OFS='\t'; done
