#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_e2e.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
for r in 1 2; do
for v in new old; do
if [ $v = old ]; then export B2_GLUE_OPS=0; else export B2_GLUE_OPS=1; fi
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$v$r.json 2> gpurun_out/${tag}_bench_$v$r.err
python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_$v$r.json'));print('$v $r', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
done
