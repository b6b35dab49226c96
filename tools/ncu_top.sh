#!/bin/bash
# developer helper (GPU box): `ncu --set full` of the five biggest kernels of a steady P-frame step of the bench workload
# (10 concurrent 1080p sessions, step 4 of tools/batch_probe.py 10 5).  Reports land in gpurun_out/<tag>_<kernel>.ncu-rep.
tag=${1:-x}
python tools/batch_probe.py 10 5 > gpurun_out/${tag}_plain.log 2>&1 || exit 1
prof() { # kernel regex, launches to skip
  local name=$(echo $1 | tr -cd "a-z_0-9")
  ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -o gpurun_out/${tag}_${name} -f \
      python tools/batch_probe.py 10 5 > gpurun_out/${tag}_ncu_${name}.log 2>&1
}
prof k_encode_rows 12
prof "^k_me\$" 9
prof k_sadmap 3
prof k_deblock_rows 4
prof k_intra_check 3
ls -la gpurun_out/${tag}_*.ncu-rep
