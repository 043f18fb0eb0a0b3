#!/bin/bash
# Per-kernel counts of the tcgen05 / TMA / cluster SASS mnemonics in the built library:
#   bash scripts/sass_evidence.sh > profiles/rNN_sass_evidence.txt
lib=plda_b200/lib/libplda_b200.so
echo "# SASS evidence (cuobjdump -sass $lib): tcgen05 / TMA / cluster mnemonics per kernel"
echo
cuobjdump -sass $lib | awk '
  /Function :/ { fn = $3 }
  { for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|UTMALDG|UTMASTG|UTMAREDG|LDTM|STTM|SYNCS|ACQBULK|UCGABAR_ARV|UCGABAR_WAIT|CGAERRBAR|REDG|RED|MAPA|DMMA)(\.|$)/) { split($i, p, "."); c[fn " " p[1]]++ } }
  END { for (k in c) { split(k, q, " "); printf "%6d\t%-12s\t%s\n", c[k], q[2], q[1] } }' | sort -t"$(printf "\t")" -k3,3 -k2,2 | c++filt | cut -c 1-170
