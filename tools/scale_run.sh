#!/bin/bash
# N-GPU bench of both dtypes on one box (run through `gpurun --gpus N -- bash tools/scale_run.sh N [extra bench flags]`)
N=$1; shift
TAG=${TAG:-r1z}
for d in f64 f32; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 200 --warmup 10 --dtype $d "$@" > gpurun_out/${TAG}_bench_${d}_${N}gpu.json 2> gpurun_out/err_${N}.log
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${d}_${N}gpu.json").read().strip().splitlines()[-1])
    print(d["dtype"], d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), round(d["e2e"]["c_abi"]["value"]), d["config"]["collective"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/err_${N}.log").read()[-1500:])
PY
done
