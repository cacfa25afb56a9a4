#!/bin/bash
# First-contact GPU script: each risky piece in its own process under `timeout` so a hang cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
CGS=""
for cg in 1 2; do
  timeout 240 python scripts/bringup_gemm.py $cg > gpurun_out/gemm_cg$cg.log 2>&1
  rc=$?
  echo "gemm cg$cg exit $rc" | tee -a gpurun_out/summary.txt
  tail -25 gpurun_out/gemm_cg$cg.log
  if [ $rc -eq 0 ] && grep -q CORRECT gpurun_out/gemm_cg$cg.log; then CGS="$CGS$cg,"; fi
done
export RLCF_TEST_CG="${CGS%,}"
echo "usable cta groups: $RLCF_TEST_CG" | tee -a gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_kernels.log
cat gpurun_out/pytest_kernels.log | tail -40
