#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "roi or pyramid" > gpurun_out/${tag}_pytest.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --config srcnn --steps 10 --warmup 3 > gpurun_out/${tag}_srcnn.json 2> gpurun_out/${tag}_srcnn.err
python -c "import json;d=json.load(open('gpurun_out/${tag}_srcnn.json'));print('srcnn', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['ms_per_step'],2))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:roi_align -s 6 -c 3 python tools/prof_r2b.py 2>&1 | grep -E "roi_align|gpu__time" | head -8
