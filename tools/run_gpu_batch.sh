#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print('bench', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks']['sm_mhz'], d['gpu_launches'])"
timeout 600 python bench.py --config srcnn --steps 10 --warmup 3 > gpurun_out/${tag}_srcnn.json 2> gpurun_out/${tag}_srcnn.err; cut -c1-200 gpurun_out/${tag}_srcnn.json; python -c "import json;d=json.load(open('gpurun_out/${tag}_srcnn.json'));print(d['clocks'])"
