#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_volume.py tests/test_gpu_properties.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print('bench', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz']); print({k:(v['ms_total'], v.get('frac_hbm')) for k,v in d['kernels'].items() if k in ('groupnorm_bwd','groupnorm_fwd','depth_head_bwd','depth_head_fwd','conv3d_c1_fwd','lift_fwd')})"
