#!/usr/bin/env bash
# ncu --set full of the three CTA-pair kernels of the graph block, each after its own plain run.  gpurun --timeout 1500 -- bash tools/ncu_pairs.sh
set -u
mkdir -p gpurun_out
timeout 120 python tools/kernel_bench.py agg_fwd knn_fwd graph_bwd --iters 1 > gpurun_out/plain_pairs.log 2>&1 || exit 1
for k in agg_fwd:agg4_tc_kernel knn_fwd:knn_pair_kernel graph_bwd:graph_bwd_pair_kernel; do
  name=${k%%:*}; kern=${k##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip 3 -c 2 -f -o gpurun_out/ncu_full_$name \
    python tools/kernel_bench.py $name --iters 1 > gpurun_out/ncu_$name.log 2>&1
  echo "$name ncu rc=$?"
done
ls -la gpurun_out/*.ncu-rep
