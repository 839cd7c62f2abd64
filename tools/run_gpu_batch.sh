#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv python tools/one_iter.py 2 > gpurun_out/${tag}_oneiter.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_oneiter.log
python tools/summarize_ncu.py gpurun_out/${tag}_launches.csv gpurun_out/${tag}_launches.txt > /dev/null; head -45 gpurun_out/${tag}_launches.txt
