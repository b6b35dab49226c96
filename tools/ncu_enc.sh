#!/bin/bash
# developer helper (GPU box): ncu --set full of the P-frame k_encode_rows sweep, 1 stream and 10 streams
tag=${1:-x}
ncu --set full --import-source on --clock-control none --kernel-name k_encode_rows --launch-skip 3 --launch-count 1 \
    -o gpurun_out/${tag}_enc_1stream -f python tools/run_small.py 1920 1080 3 > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name k_encode_rows --launch-skip 3 --launch-count 1 \
    -o gpurun_out/${tag}_enc_10stream -f python tools/batch_probe.py 10 2 > gpurun_out/${tag}_ncu10.log 2>&1
echo done
