#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; head -c 200 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
echo "ref rc=$?"; head -c 400 gpurun_out/${tag}_bench_ref.json
