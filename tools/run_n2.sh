mkdir -p gpurun_out
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n${N}_bench.log 2> gpurun_out/n${N}_bench.err; echo "bench N=$N rc=$?"; tail -1 gpurun_out/n${N}_bench.log; tail -5 gpurun_out/n${N}_bench.err
