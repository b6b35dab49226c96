#!/bin/sh
# Regenerates h264-lab_b200/csrc/h264_denoise_tab.h (needs /root/reference).
set -e
REF=${REF:-/root/reference/src}
T=$(mktemp -d)
gcc -w -I"$REF" -o "$T/gen" "$(dirname "$0")/gen_denoise_table.c" -lm
"$T/gen" > "$(dirname "$0")/../h264-lab_b200/csrc/h264_denoise_tab.h"
rm -rf "$T"
