#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_volume.py -q -m gpu -k "conv3d" > gpurun_out/${tag}_pytest_dc.log 2>&1
rc=$?; echo "dc pair conv tests rc=$rc"; tail -6 gpurun_out/${tag}_pytest_dc.log
if [ $rc -ne 0 ]; then export B2_CONV_DC_PAIR=0; echo "FALLING BACK to B2_CONV_DC_PAIR=0"; fi
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_round2.py -q -m gpu --maxfail=10 > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -8 gpurun_out/${tag}_pytest.log
timeout 300 python tools/bench_conv3d_shapes.py > gpurun_out/${tag}_conv3d_shapes.log 2>&1; cat gpurun_out/${tag}_conv3d_shapes.log
B2_CONV_DC_PAIR=0 timeout 300 python tools/bench_conv3d_shapes.py > gpurun_out/${tag}_conv3d_shapes_nopair.log 2>&1; cat gpurun_out/${tag}_conv3d_shapes_nopair.log
timeout 300 python tools/bench_lift.py > gpurun_out/${tag}_bench_lift.log 2>&1; cat gpurun_out/${tag}_bench_lift.log
timeout 300 python tools/measure_tf32_peak.py > gpurun_out/${tag}_tf32.log 2>&1; cat gpurun_out/${tag}_tf32.log
timeout 600 python tests/diag/diag_fullsize.py > gpurun_out/${tag}_diag_fullsize.log 2>&1; tail -4 gpurun_out/${tag}_diag_fullsize.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 200 gpurun_out/${tag}_bench.json
