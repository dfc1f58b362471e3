#!/bin/bash
# usage: tools/sass_count.sh <mangled-name-substring> : per-opcode instruction histogram of one kernel of the library
LIB=neural_admixture_b200/csrc/libnadm_b200.so
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function :/ {on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\/$//'
