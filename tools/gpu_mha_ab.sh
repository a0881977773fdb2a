#!/bin/bash
# A/B of the ViT-B/16 attention kernel variants (DFD_MHA_VARIANT): correctness + time per launch
mkdir -p gpurun_out
: > gpurun_out/mha_ab.log
for v in ${VARIANTS:-1 4}; do
  echo "== variant $v" >> gpurun_out/mha_ab.log
  DFD_MHA_VARIANT=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "mha" 2>&1 | tail -3 >> gpurun_out/mha_ab.log
  for r in 1 2; do DFD_MHA_VARIANT=$v timeout 120 python tools/prof_kernel.py mha 50 >> gpurun_out/mha_ab.log 2>&1; done
done
cat gpurun_out/mha_ab.log
