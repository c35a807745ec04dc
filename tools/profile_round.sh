#!/bin/bash
# Round-end profiling recipe (run on the GPU box through gpurun; every ncu pass follows a plain run of the same
# command that exited 0).  Outputs under gpurun_out/, copied to profiles/ by hand.
set -u
T=${1:-r2}
O=gpurun_out
python tools/run_step.py --steps 2 || exit 1
python tools/run_step.py --steps 2 --dtype f32 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"mma_kernel|kuu_fast|kdiag_bwd_saved" -s 6 -c 6 \
    -o $O/${T}_f64 python tools/run_step.py --steps 3 > $O/${T}_ncu_f64.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"kuf_fwd_kernel|kdiag_fwd_kernel|kuf_bwd_kernel|rowsum_kernel|kuf_dz_kernel|kuu_fast|kdiag_bwd_saved" -s 6 -c 6 \
    -o $O/${T}_f32 python tools/run_step.py --steps 3 --dtype f32 > $O/${T}_ncu_f32.log 2>&1
for d in f64 f32; do
  ncu -i $O/${T}_$d.ncu-rep --page raw --csv > $O/${T}_ncu_full_${d}_raw.csv 2>/dev/null
  for k in 0 1 3; do python tools/ncu_source_hot.py $O/${T}_$d.ncu-rep $k 25 > $O/${T}_hot_${d}_$k.txt 2>&1; done
done
rm -f $O/${T}_f64.ncu-rep $O/${T}_f32.ncu-rep
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench_f64.csv \
    python bench.py --steps 2 --warmup 1 > $O/${T}_ncu_launch_f64.log 2>&1
python bench.py --steps 2 --warmup 1 --dtype f32 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_bench_f32.csv \
    python bench.py --steps 2 --warmup 1 --dtype f32 > $O/${T}_ncu_launch_f32.log 2>&1
ls -la $O | tail -20
