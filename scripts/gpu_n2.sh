#!/usr/bin/env bash
# Two-GPU visit: the NCCL/peer-memory sharded tests, then the bench as the driver launches it for N = 2.
set -u
mkdir -p gpurun_out
TAG="${1:-n2}"
N="${2:-2}"
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_sharded_$TAG.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.json
grep "bench rank 0" gpurun_out/bench_$TAG.err | tail -12
