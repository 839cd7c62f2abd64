#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s > gpurun_out/${tag}_fullsize.log 2>&1
echo "rc=$?"; grep -n "CONFIG\|free-running pair\|passed\|failed" gpurun_out/${tag}_fullsize.log | cut -c1-400
