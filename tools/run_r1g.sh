mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1g_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r1g_pytest.log)"
python tools/kernel_bench.py 128 gemm768 > gpurun_out/r1g_gemm768.log 2>&1; echo "kb rc=$?"; cat gpurun_out/r1g_gemm768.log
python bench.py > gpurun_out/r1g_bench.log 2> gpurun_out/r1g_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/r1g_bench.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/r1g_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r1g_ncu.log 2>&1; echo "ncu rc=$?"
