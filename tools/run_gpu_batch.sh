#!/bin/bash
# usage: tools/run_gpu_batch.sh <tag>   (runs on the GPU box through gpurun; logs into gpurun_out/)
tag=${1:-b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_conv2d.py -q -m gpu -s > gpurun_out/${tag}_conv2d.log 2>&1
echo "conv2d tests rc=$?"; tail -3 gpurun_out/${tag}_conv2d.log
timeout 300 python tests/diag/diag_pgd_tiny.py > gpurun_out/${tag}_diag_tiny.log 2>&1
echo "diag tiny rc=$?"; tail -12 gpurun_out/${tag}_diag_tiny.log
timeout 1200 python -m pytest tests -q -m gpu --maxfail=12 --deselect tests/test_gpu_conv2d.py > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu suite rc=$?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python tools/bench_2d.py > gpurun_out/${tag}_bench2d.log 2>&1
echo "bench2d rc=$?"; tail -8 gpurun_out/${tag}_bench2d.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/${tag}_bench.json
