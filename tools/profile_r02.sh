#!/bin/bash
# Round-2 GPU evidence in one gpurun call: (0) GPU tests, (1) the plain bench line, (2) the reference arm (one step),
# (3) ncu launch list of ONE evaluation of the same command, (4) ncu --set full of the dataflow Cholesky launches of that
# evaluation.  usage (under gpurun): tools/profile_r02.sh <tag> [skip-tests]
tag=${1:-r02}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
  echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
fi
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
echo "reference rc=$?"; cat gpurun_out/bench_ref_$tag.json
CMD="python bench.py --steps 1 --warmup 3 --evals-per-step 1 --lanes 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/bench_plain_$tag.json 2> gpurun_out/bench_plain_$tag.err || { echo "plain run failed"; tail -5 gpurun_out/bench_plain_$tag.err; exit 1; }
L=$(python -c "import json;print(json.loads(open('gpurun_out/bench_plain_$tag.json').read().strip().splitlines()[-1])['gpu_launches'])")
echo "launches per evaluation: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(?!dmma_peak)' -s $((3*L)) -c $L --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_launch_$tag.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:potrf_dataflow -s 18 -c 6 -f -o gpurun_out/potrf_$tag $CMD > gpurun_out/ncu_potrf_$tag.log 2>&1
echo "potrf capture rc=$?"
ls -la gpurun_out/ | tail -14
