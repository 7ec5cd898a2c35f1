mkdir -p gpurun_out
T=${1:-dw}
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dw3x3" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)"
echo "== MINB=3"; python tools/kernel_bench.py 128 dw > gpurun_out/${T}_kb3.log 2>&1; echo "kb rc=$?"; cat gpurun_out/${T}_kb3.log
echo "== MINB=2"; XCP_DW_FWD_MINB=2 python tools/kernel_bench.py 128 dw > gpurun_out/${T}_kb2.log 2>&1; echo "kb rc=$?"; grep fwd gpurun_out/${T}_kb2.log
