#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) full captures of the dominant GEMM launches
# (LAUUM / SYRK shaped, 128x128 configuration) and of the dataflow Cholesky.  usage (under gpurun): tools/profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain_$tag.json 2> gpurun_out/bench_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/bench_plain_$tag.err; exit 1; }
cat gpurun_out/bench_plain_$tag.json
L=$(python -c "import json;print(json.loads(open('gpurun_out/bench_plain_$tag.json').read().strip().splitlines()[-1])['gpu_launches']//2)")
echo "launches per step: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+2)) -c $L --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
# TN (0,0) = LAUUM N=8192 then n=4096; NT (1,1) = gradient SYRK n^2 N
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'6, .bool.0, .bool.0' -c 1 -o gpurun_out/gemm_lauum_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "lauum capture rc=$?"
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'6, .bool.1, .bool.1' -c 1 -o gpurun_out/gemm_syrk_$tag $CMD >> gpurun_out/ncu_full_$tag.log 2>&1
echo "syrk capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -c 1 -o gpurun_out/potrf_$tag $CMD > gpurun_out/ncu_potrf_$tag.log 2>&1
echo "potrf capture rc=$?"
ls -la gpurun_out/ | tail -12
