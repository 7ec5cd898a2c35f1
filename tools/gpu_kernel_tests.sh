#!/bin/bash
# Runs the kernel-level GPU tests group by group in separate processes (a faulting kernel poisons its CUDA
# context, so isolation keeps the other groups' verdicts meaningful).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in gemm_tn_bf16 gemm_tn_f32 gemm_wgrad conv3x3 stem_conv1 dw3x3 bn_finalize pool_add bn_bwd_modes layout lstm head_linear arcface adam; do
  timeout -k 5 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" --timeout 120 -x --no-header -p no:cacheprovider > gpurun_out/k_$grp.log 2>&1
  echo "$grp rc=$? $(tail -1 gpurun_out/k_$grp.log)"
done
