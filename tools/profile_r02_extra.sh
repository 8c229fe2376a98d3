#!/bin/bash
# ncu --set full captures beyond the headline step: a Pade product of expm at N = 8192 (the flop-bound GEMM), the Jacobi pivot
# kernel and the batched small-n evaluator.  usage (under gpurun): tools/profile_r02_extra.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
python tools/expm_once.py > gpurun_out/expm_once_$tag.log 2>&1 || { echo "expm_once failed"; tail -5 gpurun_out/expm_once_$tag.log; }
tail -1 gpurun_out/expm_once_$tag.log
# second build: skip the launches of the first (the build is ~26 launches; the first big GEMM after launch 30 is A^2 of build 2)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_f64_kernel<2' -s 9 -c 1 -f -o gpurun_out/gemm_expm_$tag python tools/expm_once.py > gpurun_out/ncu_gemm_expm_$tag.log 2>&1
echo "gemm expm capture rc=$?"
python tools/spectral_once.py > gpurun_out/spectral_once_$tag.log 2>&1 || { echo "spectral_once failed"; tail -5 gpurun_out/spectral_once_$tag.log; }
tail -1 gpurun_out/spectral_once_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spectral_batch|jacobi64' -s 40 -c 3 -f -o gpurun_out/spectral_$tag python tools/spectral_once.py > gpurun_out/ncu_spectral_$tag.log 2>&1
echo "spectral capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spectral_batch' -c 1 -f -o gpurun_out/spectral_batch_$tag python tools/spectral_once.py > gpurun_out/ncu_spectral_batch_$tag.log 2>&1
echo "spectral batch capture rc=$?"
ls -la gpurun_out | tail -8
