#!/bin/bash
# usage: tools/run_gpu_batch.sh <tag>   (runs on the GPU box through gpurun; logs into gpurun_out/)
tag=${1:-b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv2d.py -q -m gpu -s > gpurun_out/${tag}_conv2d.log 2>&1
echo "conv2d tests rc=$?"; tail -3 gpurun_out/${tag}_conv2d.log; grep "rel. error" gpurun_out/${tag}_conv2d.log | head -2
timeout 300 python tests/diag/diag_pgd_tiny2.py > gpurun_out/${tag}_diag_tiny2.log 2>&1
echo "diag tiny2 rc=$?"; tail -30 gpurun_out/${tag}_diag_tiny2.log
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -s > gpurun_out/${tag}_round2.log 2>&1
echo "round2 tests rc=$?"; tail -15 gpurun_out/${tag}_round2.log
timeout 900 python tests/diag/diag_fullsize.py > gpurun_out/${tag}_diag_fullsize.log 2>&1
echo "diag fullsize rc=$?"; tail -6 gpurun_out/${tag}_diag_fullsize.log
