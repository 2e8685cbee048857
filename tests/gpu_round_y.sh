#!/bin/bash
# cooperative launch of gn_persistent_kernel: with / without PDL on that launch, vs the plain launch
mkdir -p gpurun_out
LR_CASE_TIMEOUT=90 timeout 300 python tests/gpu_diag_ops.py --only gn_bigmean_twopass,gn 2>&1 | tail -1
LR_GN_COOP_PDL=1 LR_CASE_TIMEOUT=90 timeout 300 python tests/gpu_diag_ops.py --only gn_bigmean_twopass 2>&1 | tail -2
for rep in 1 2; do
  echo "coop:      $(timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)"
  echo "coop+pdl:  $(LR_GN_COOP_PDL=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)"
  echo "plain:     $(LR_GN_NO_COOP=1 timeout 300 python tests/gpu_time_forward.py 40 2>&1 | tail -1)"
done
