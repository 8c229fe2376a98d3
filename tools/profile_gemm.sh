#!/bin/bash
# full ncu capture of the dominant GEMM launch of a bench step (128x128 configuration): the 8th big launch = top level of the
# triangular inverse (W21 = -W22 T: 1024 tiles with triangular k-ranges).  usage (under gpurun): tools/profile_gemm.sh <tag>
tag=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain2_$tag.json 2> gpurun_out/bench_plain2_$tag.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'6, .bool.1, .bool.1' --launch-skip 7 -c 1 -f -o gpurun_out/gemm_trtri_$tag $CMD > gpurun_out/ncu_trtri_$tag.log 2>&1
echo "trtri capture rc=$?"
