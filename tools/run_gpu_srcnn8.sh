#!/bin/bash
tag=${1:-m}; n=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $n --config srcnn --steps 10 --warmup 3 > gpurun_out/${tag}_bench_srcnn_n${n}.json 2> gpurun_out/${tag}_bench_srcnn_n${n}.err
echo "srcnn n=$n rc=$?"; python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_srcnn_n${n}.json'));print(round(d['value'],2), 'e2e', round(d['e2e']['value'],2), round(d['ms_per_step'],2), d['clocks'])"
