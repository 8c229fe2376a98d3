#!/bin/bash
# usage: tools/run_stages.sh stage1 stage2 ...   (each stage in its own process, own timeout, own log)
mkdir -p gpurun_out
for s in "$@"; do
  echo "##### stage $s" 
  timeout 150 python tools/gpu_check.py $s > gpurun_out/stage_$s.log 2>&1
  echo "rc=$?"
  tail -n 60 gpurun_out/stage_$s.log
done
