#!/bin/bash
# Round-end multi-GPU pass on one 8-GPU box (run through `gpurun --gpus 8 -- bash tools/scale_all.sh TAG`):
# the 2-rank GPU parity tests, then the default bench line at N = 8, 4, 2 (weak + strong + grad_check + configs + predict).
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_multirank.log
for N in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
  echo "N=$N rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_${N}gpu.err
done
cat gpurun_out/${TAG}_pytest_multirank.log
