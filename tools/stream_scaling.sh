#!/bin/bash
# developer helper (GPU box): macroblock-sweep time vs number of concurrent streams
for n in 1 2 3 5 7 10 14; do echo "== streams $n"; timeout 300 python tools/batch_probe.py $n 5 2>&1 | tail -2 | cut -c1-64; done > gpurun_out/stream_scaling.log 2>&1
echo done
