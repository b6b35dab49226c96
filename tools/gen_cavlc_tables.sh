#!/bin/sh
# Regenerates h264-lab_b200/csrc/h264_cavlc_tables.h (needs /root/reference).
set -e
REF=${REF:-/root/reference/src}
T=$(mktemp -d)
gcc -w -I"$REF" -o "$T/gen" "$(dirname "$0")/gen_cavlc_tables.c" -lm
"$T/gen" > "$(dirname "$0")/../h264-lab_b200/csrc/h264_cavlc_tables.h"
rm -rf "$T"
