mkdir -p gpurun_out
T=${1:-full}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${T}_pytest.log)"
python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/${T}_bench.log | cut -c1-330; tail -3 gpurun_out/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 2400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1; echo "ncu rc=$?"
