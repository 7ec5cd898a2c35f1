mkdir -p gpurun_out
T=${1:-g}
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm or conv3x3 or stem" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)"
python tools/kernel_bench.py 128 gemm > gpurun_out/${T}_kb.log 2>&1; echo "kb rc=$?"; cat gpurun_out/${T}_kb.log
