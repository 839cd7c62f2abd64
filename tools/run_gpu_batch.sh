#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/one_iter.py 3 > gpurun_out/${tag}_one_iter.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${tag}_one_iter.log; wc -l gpurun_out/${tag}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lift_fwd|grid_sample_bwd|conv2d_halo|conv3d_dc2|conv3d_g2|cost_volume_bwd_row|pgd_update" -s 9 -c 9 -o gpurun_out/${tag}_prof -f python tools/prof_r2.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_ncu.log
