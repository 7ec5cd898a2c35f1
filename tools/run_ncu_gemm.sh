mkdir -p gpurun_out
T=${1:-ncug}
python tools/ncu_kernels.py 256 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|dw3x3|bnbwd' -s 7 -c 7 -o gpurun_out/${T}_prof python tools/ncu_kernels.py 256 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${T}_ncu.log
