#!/bin/bash
# runs tools/gemm_race.py over the kernel variants; usage (under gpurun): tools/gemm_race_matrix.sh [iters]
mkdir -p gpurun_out
it=${1:-150}
export GEMM_RACE_LIB=${2:-}
for dbg in 0 0x100 0x200 0x400 0x800 0x80; do
  GPNET_B200_GEMM_DEBUG=$dbg ITERS=$it timeout 300 python tools/gemm_race.py 2>&1 | tee -a gpurun_out/gemm_race.log
done
GPNET_B200_GEMM_DEBUG=0 OFF=128 ITERS=$it timeout 300 python tools/gemm_race.py 2>&1 | tee -a gpurun_out/gemm_race.log
