mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r1h_pytest.log)"
python tools/kernel_bench.py 128 gemm768 > gpurun_out/r1h_gemm768.log 2>&1; echo "kb rc=$?"; grep -E "K=768 N=768|K=728 N=728" gpurun_out/r1h_gemm768.log
python bench.py --no-cpu-baseline > gpurun_out/r1h_bench.log 2> gpurun_out/r1h_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r1h_bench.log; tail -3 gpurun_out/r1h_bench.err
