#!/bin/bash
# compute-sanitizer passes over the kernel-level tests (run through gpurun; summaries -> gpurun_out/, keep under profiles/).
#   tools/sanitize.sh <tag>
# PYTORCH_NO_CUDA_MEMORY_CACHING=1 gives every tensor its own cudaMalloc so that memcheck sees out-of-bounds accesses that the
# caching allocator's pools would hide.  racecheck / synccheck look at the hand-rolled mbarrier / TMA / cluster-DSMEM protocols.
tag=${1:-san}
out=gpurun_out
mkdir -p $out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SEL_MEM="tests/test_kernel_twins_gpu.py tests/test_kernels_gpu.py"
SEL_RACE="tests/test_kernel_twins_gpu.py tests/test_kernels_gpu.py::test_lstm_fwd_bwd tests/test_kernels_gpu.py::test_gemm_tn_bf16_stats tests/test_kernels_gpu.py::test_gemm_wgrad tests/test_kernels_gpu.py::test_conv3x3_implicit_gemm tests/test_kernels_gpu.py::test_dw3x3_bwd_residual_adds tests/test_kernels_gpu.py::test_head_linear_bce"
run() {   # tool, timeout, selection...
    tool=$1; lim=$2; shift 2
    start=$(date +%s)
    timeout $lim compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 20 --log-file $out/${tag}_$tool.raw \
        python -m pytest "$@" -q -x -p no:cacheprovider > $out/${tag}_$tool.pytest 2>&1
    rc=$?
    end=$(date +%s)
    { echo "== compute-sanitizer --tool $tool (rc=$rc, $((end-start)) s; rc 0 = tests passed and no sanitizer error, 99 = sanitizer error, 124 = timeout)";
      echo "selection: $@";
      tail -3 $out/${tag}_$tool.pytest;
      grep -a "ERROR SUMMARY\|RACECHECK SUMMARY\|hazard\|Invalid\|error" $out/${tag}_$tool.raw | sort | uniq -c | head -20; } > $out/${tag}_$tool.summary
    cat $out/${tag}_$tool.summary
}
run memcheck 1500 $SEL_MEM
run synccheck 900 $SEL_RACE
run racecheck 1500 $SEL_RACE
