#!/bin/bash
# developer helper (GPU box): `ncu --set full` evidence for tools/ncu_summarise.py.  gpurun brings back at most
# 64 MiB per call, so the two halves are separate calls:
#   tools/ncu_all.sh <tag> enc   : the k_encode_rows sweep alone with source correlation (1 stream and 10 streams)
#   tools/ncu_all.sh <tag> step  : EVERY kernel launch of one P-frame step of the bench workload (10 concurrent 1080p frames)
tag=${1:-x}; what=${2:-enc}
if [ "$what" = step ]; then
  # a step launches 25 kernels (init + 3 + 2 x (3 repair rounds + 4) + 6 + gather); step 3 is a P frame with a predicted trajectory
  ncu --set full --clock-control none --launch-skip 75 --launch-count 25 \
      -o gpurun_out/${tag}_step_allkernels -f python tools/batch_probe.py 10 4 > gpurun_out/${tag}_ncu_all.log 2>&1
else
  ncu --set full --import-source on --clock-control none --kernel-name k_encode_rows --launch-skip 9 --launch-count 1 \
      -o gpurun_out/${tag}_enc_10stream -f python tools/batch_probe.py 10 4 > gpurun_out/${tag}_ncu10.log 2>&1
  ncu --set full --import-source on --clock-control none --kernel-name k_encode_rows --launch-skip 9 --launch-count 1 \
      -o gpurun_out/${tag}_enc_1stream -f python tools/batch_probe.py 1 4 > gpurun_out/${tag}_ncu1.log 2>&1
fi
ls -la gpurun_out/${tag}_*.ncu-rep
echo done
