mkdir -p gpurun_out
T=${1:-dw}
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dw3x3" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)"
python tools/kernel_bench.py 128 dw > gpurun_out/${T}_kb.log 2>&1; echo "kb rc=$?"; cat gpurun_out/${T}_kb.log
python tools/ncu_kernels.py 128 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw3x3' -s 2 -c 4 -o gpurun_out/${T}_prof python tools/ncu_kernels.py 128 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${T}_ncu.log
