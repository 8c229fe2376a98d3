#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench step, (2) full capture of the dataflow Cholesky
# (the dominant GEMM launches are captured by tools/profile_gemm.sh).  usage (under gpurun): tools/profile.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/bench_plain_$tag.json 2> gpurun_out/bench_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/bench_plain_$tag.err; exit 1; }
cat gpurun_out/bench_plain_$tag.json
L=$(python -c "import json;print(json.loads(open('gpurun_out/bench_plain_$tag.json').read().strip().splitlines()[-1])['gpu_launches']//2)")
echo "launches per step: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(?!dmma_peak)' -s $((3*L)) -c $L --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -c 1 -f -o gpurun_out/potrf_$tag $CMD > gpurun_out/ncu_potrf_$tag.log 2>&1
echo "potrf capture rc=$?"
ls -la gpurun_out/ | tail -12
