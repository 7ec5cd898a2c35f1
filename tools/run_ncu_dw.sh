mkdir -p gpurun_out
T=${1:-ncudw}
python tools/ncu_dw.py 32 147 128 > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dw3x3' -s 3 -c 3 -o gpurun_out/${T}_prof python tools/ncu_dw.py 32 147 128 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${T}_ncu.log
