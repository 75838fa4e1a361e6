#!/bin/bash
# SASS opcode histogram per cubin of libpgm_b200.so (evidence of what the kernels use: UBLKCP / SYNCS = bulk-TMA program
# staging, D* = FP64 vector math, no UTC*MMA / HMMA: no tensor-core path).  Usage: tools/sass_opcodes.sh > profiles/sass_opcodes.txt
LIB=${1:-power-grid-model_b200/libpgm_b200.so}
TMP=$(mktemp -d)
(cd "$TMP" && cuobjdump -xelf all "$OLDPWD/$LIB" > /dev/null)
for f in "$TMP"/*.cubin; do
  echo "== $(basename "$f")"
  cuobjdump -sass "$f" | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed -E 's/\..*//; s/;$//' \
    | sort | uniq -c | sort -rn | awk '{printf "%s:%s ", $2, $1} END {print ""}' | fold -w 160 -s
done
rm -rf "$TMP"
