#!/bin/bash
# full ncu capture of the GEMM launch of a default bench step (128x128 configuration): Z = W_oo^T W_oo, 528 lower tiles with
# triangular k-ranges, mirrored store.  usage (under gpurun): tools/profile_gemm.sh <tag>
tag=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain2_$tag.json 2> gpurun_out/bench_plain2_$tag.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'6, .bool.1, .bool.1' --launch-skip 2 -c 1 -f -o gpurun_out/gemm_lauum_$tag $CMD > gpurun_out/ncu_lauum_$tag.log 2>&1
echo "lauum capture rc=$?"
