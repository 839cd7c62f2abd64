#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_volume.py tests/test_gpu_properties.py -q -m gpu --maxfail=10 > gpurun_out/${tag}_pytest_conv.log 2>&1
rc=$?; echo "conv tests rc=$rc"; tail -6 gpurun_out/${tag}_pytest_conv.log
timeout 300 python tools/bench_conv3d_shapes.py > gpurun_out/${tag}_conv3d_shapes.log 2>&1; cat gpurun_out/${tag}_conv3d_shapes.log
timeout 900 python -m pytest tests -q -m gpu --maxfail=10 --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_volume.py --deselect tests/test_gpu_properties.py > gpurun_out/${tag}_pytest.log 2>&1
echo "other gpu tests rc=$?"; tail -8 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 200 gpurun_out/${tag}_bench.json
