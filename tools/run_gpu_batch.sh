#!/bin/bash
# usage: tools/run_gpu_batch.sh <tag>   (runs on the GPU box through gpurun; logs into gpurun_out/)
tag=${1:-b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt
timeout 600 python -m pytest tests/test_gpu_conv2d.py -q -m gpu -s > gpurun_out/${tag}_conv2d.log 2>&1
echo "conv2d tests rc=$?"; tail -3 gpurun_out/${tag}_conv2d.log; grep "^conv2d\|rel. error" gpurun_out/${tag}_conv2d.log | head -40
timeout 1500 python -m pytest tests -q -m gpu --maxfail=20 --deselect tests/test_gpu_conv2d.py --deselect tests/test_gpu_fullsize.py > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu suite rc=$?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 1800 python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s > gpurun_out/${tag}_fullsize.log 2>&1
echo "fullsize rc=$?"; grep "CONFIG\|iter \|passed\|failed\|Error" gpurun_out/${tag}_fullsize.log | tail -40
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/${tag}_bench.json
