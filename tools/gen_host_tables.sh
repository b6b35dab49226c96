#!/bin/sh
# Regenerates h264-lab_b200/host/h264e_tables_gen.h (needs /root/reference).
set -e
REF=${REF:-/root/reference/src}
T=$(mktemp -d)
gcc -w -I"$REF" -o "$T/gen" "$(dirname "$0")/gen_host_tables.c" -lm
"$T/gen" > "$(dirname "$0")/../h264-lab_b200/host/h264e_tables_gen.h"
rm -rf "$T"
