mkdir -p gpurun_out
T=${1:-q}
python -m pytest tests/test_kernels_gpu.py tests/test_optim_gpu.py -m gpu -x -q -k "gemm or conv3x3 or adam or optim" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)"
python tools/kernel_bench.py 128 gemm768 > gpurun_out/${T}_gemm768.log 2>&1; echo "kb rc=$?"; grep -E "K=768" gpurun_out/${T}_gemm768.log
python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${T}_bench.log | cut -c1-400; tail -3 gpurun_out/${T}_bench.err
