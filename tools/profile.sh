#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) full captures of the GEMM launches of one step
# and of the dataflow Cholesky.  usage (under gpurun): tools/profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain_$tag.json 2> gpurun_out/bench_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/bench_plain_$tag.err; exit 1; }
cat gpurun_out/bench_plain_$tag.json
L=$(python -c "import json;print(json.loads(open('gpurun_out/bench_plain_$tag.json').read().strip().splitlines()[-1])['gpu_launches']//2)")
echo "launches per step: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3*L+2)) -c $L --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -c 30 -o gpurun_out/gemm_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "gemm capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -c 2 -o gpurun_out/potrf_$tag $CMD > gpurun_out/ncu_potrf_$tag.log 2>&1
echo "potrf capture rc=$?"
ls -la gpurun_out/ | tail -12
