#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "roi or srcnn or stereo_rcnn or pyramid" > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --config srcnn --steps 10 --warmup 3 > gpurun_out/${tag}_srcnn.json 2> gpurun_out/${tag}_srcnn.err
python -c "import json;d=json.load(open('gpurun_out/${tag}_srcnn.json'));print('srcnn', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['ms_per_step'],2), d['clocks'])"
