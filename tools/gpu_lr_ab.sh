#!/bin/bash
# A/B of the fp16-plane product kernel's tile / split configuration (kernel_times.py lowrank, r = 32 and 64).
for cfg in "0 2" "1 2" "2 2" "3 1" "0 3" "0 1" "1 3" "0 4"; do
  set -- $cfg
  echo "== CF_LR_HB_TILE=$1 CF_LR_CTAS_PER_SM=$2"
  for r in 32; do
    CF_LR_HB_TILE=$1 CF_LR_CTAS_PER_SM=$2 timeout 100 python tools/kernel_times.py lowrank --rank $r 2>&1 | grep -E "k_lr_gemm|k_lr_orth|sum of" | cut -c1-60,100-140
  done
done
