#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 600 python tests/diag/diag_gn2d.py > gpurun_out/${tag}_diag.log 2>&1; head -12 gpurun_out/${tag}_diag.log
timeout 900 python -m pytest tests/test_gpu_conv2d.py tests/test_gpu_round2.py tests/test_gpu_e2e.py tests/test_gpu_properties.py tests/test_gpu_volume.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
