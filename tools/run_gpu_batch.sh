#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv2d.py tests/test_gpu_volume.py tests/test_gpu_round2.py tests/test_gpu_e2e.py -q -m gpu --maxfail=20 > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python tools/bench_lift.py > gpurun_out/${tag}_bench_lift.log 2>&1
B2_GS_BWD_WIDE=0 timeout 300 python tools/bench_lift.py > gpurun_out/${tag}_bench_lift_narrow.log 2>&1
echo "bench_lift rc=$?"; cat gpurun_out/${tag}_bench_lift.log gpurun_out/${tag}_bench_lift_narrow.log | tail -8
timeout 600 python tools/bench_2d.py > gpurun_out/${tag}_bench2d.log 2>&1
echo "bench2d rc=$?"; cat gpurun_out/${tag}_bench2d.log | tail -20
B2_CONV2D_HALO=0 timeout 600 python tools/bench_2d.py > gpurun_out/${tag}_bench2d_nohalo.log 2>&1
echo "bench2d nohalo rc=$?"; cat gpurun_out/${tag}_bench2d_nohalo.log | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 300 gpurun_out/${tag}_bench.json
