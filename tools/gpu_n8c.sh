O=gpurun_out; T=${1:-n8}; N=${2:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $R bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_n${N}_$T.log 2> $O/r02_bench_n${N}_$T.err; tail -c 300 $O/r02_bench_n${N}_$T.log; tail -3 $O/r02_bench_n${N}_$T.err
