mkdir -p gpurun_out
T=${1:-ew}
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dw3x3 or bn_bwd or pool" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)"
python tools/kernel_bench.py 128 ew > gpurun_out/${T}_kb.log 2>&1; echo "kb rc=$?"; cat gpurun_out/${T}_kb.log
python tools/kernel_bench.py 128 dw 2>&1 | grep -E "19x19"
