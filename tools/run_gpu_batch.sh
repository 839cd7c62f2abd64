#!/bin/bash
# final validation batch on one B200: GPU test suite, smoke, default bench, reference arm, configs 4 and 5
tag=${1:-b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print('bench', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'], d['gpu_launches'])"
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py --config patch --steps 50 --warmup 4 > gpurun_out/${tag}_patch.json 2> gpurun_out/${tag}_patch.err; python -c "import json;d=json.load(open('gpurun_out/${tag}_patch.json'));print('patch', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'])"
timeout 600 python bench.py --config srcnn --steps 10 --warmup 3 > gpurun_out/${tag}_srcnn.json 2> gpurun_out/${tag}_srcnn.err; python -c "import json;d=json.load(open('gpurun_out/${tag}_srcnn.json'));print('srcnn', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['clocks'])"
