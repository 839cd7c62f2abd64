#!/bin/bash
# usage: tools/run_gpu_multi.sh <tag> <ngpus>
tag=${1:-m}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "ranks" > gpurun_out/${tag}_pytest_ranks.log 2>&1
echo "sharding invariance rc=$?"; tail -3 gpurun_out/${tag}_pytest_ranks.log
timeout 900 bash -c "$(declare -f run); n=$n; run 29541 bench.py --gpus $n --config patch --steps 50 --warmup 4" > gpurun_out/${tag}_bench_patch_n${n}.json 2> gpurun_out/${tag}_bench_patch_n${n}.err
echo "patch n=$n rc=$?"; head -c 400 gpurun_out/${tag}_bench_patch_n${n}.json; tail -3 gpurun_out/${tag}_bench_patch_n${n}.err
timeout 900 bash -c "$(declare -f run); n=$n; run 29542 bench.py --gpus $n --config srcnn --steps 10 --warmup 3" > gpurun_out/${tag}_bench_srcnn_n${n}.json 2> gpurun_out/${tag}_bench_srcnn_n${n}.err
echo "srcnn n=$n rc=$?"; head -c 400 gpurun_out/${tag}_bench_srcnn_n${n}.json; tail -3 gpurun_out/${tag}_bench_srcnn_n${n}.err
timeout 900 bash -c "$(declare -f run); n=$n; run 29543 bench.py --gpus $n --steps 10 --warmup 3" > gpurun_out/${tag}_bench_pgd_n${n}.json 2> gpurun_out/${tag}_bench_pgd_n${n}.err
echo "pgd n=$n rc=$?"; head -c 300 gpurun_out/${tag}_bench_pgd_n${n}.json; tail -3 gpurun_out/${tag}_bench_pgd_n${n}.err
if [ "$3" == "single" ]; then
  CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --config patch --steps 50 --warmup 4 > gpurun_out/${tag}_bench_patch_n1.json 2> gpurun_out/${tag}_bench_patch_n1.err
  echo "patch n=1 rc=$?"; head -c 300 gpurun_out/${tag}_bench_patch_n1.json
  CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --config srcnn --steps 10 --warmup 3 > gpurun_out/${tag}_bench_srcnn_n1.json 2> gpurun_out/${tag}_bench_srcnn_n1.err
  echo "srcnn n=1 rc=$?"; head -c 300 gpurun_out/${tag}_bench_srcnn_n1.json
fi
