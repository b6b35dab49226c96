#!/bin/bash
# developer helper (GPU box): per-launch durations of a short bench run
tag=${1:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    env H264B200_NO_LIVE_PEAK=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
echo done
