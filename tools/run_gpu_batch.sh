#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 300 python tools/bench_conv3d_shapes.py > gpurun_out/${tag}_conv3d_shapes.log 2>&1; cat gpurun_out/${tag}_conv3d_shapes.log
timeout 600 python -m pytest tests/test_gpu_volume.py tests/test_gpu_conv2d.py tests/test_gpu_round2.py -q -m gpu --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -8 gpurun_out/${tag}_pytest.log
timeout 300 python tools/bench_lift.py > gpurun_out/${tag}_bench_lift.log 2>&1; cat gpurun_out/${tag}_bench_lift.log
