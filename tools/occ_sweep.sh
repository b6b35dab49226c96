#!/bin/bash
# developer helper (GPU box): sweep-0 time vs CTAs resident per SM
for sm in 0 30000 42000 60000 100000; do
  echo "== extra smem $sm"; H264B200_ENC_SMEM=$sm python tools/batch_probe.py 10 5 2>&1 | tail -2
done > gpurun_out/occ_sweep.log 2>&1
echo done
