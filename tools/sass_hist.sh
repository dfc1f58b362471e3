#!/bin/bash
# usage: tools/sass_hist.sh <lib.so> <mangled-name-substring>... : opcode histogram (top 25) of each named kernel.
# Evidence for "which instructions the hot kernels are made of": UTCIMMA/UTCHMMA = tcgen05.mma kind::i8 / kind::f16,
# LDTM/STTM = tcgen05.ld/st, LDGSTS = cp.async, UTMALDG/UBLKCP would be TMA (none in the library: see DESIGN 3.1).
LIB=$1; shift
for pat in "$@"; do
  echo "== $pat"
  cuobjdump -sass "$LIB" | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on' \
    | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' \
    | awk '{split($1, a, "."); n[a[1]]++; t++} END {for (k in n) printf "%7d %s\n", n[k], k; printf "%7d TOTAL\n", t}' | sort -rn | head -45
done
