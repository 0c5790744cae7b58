# one 8-GPU box: end-to-end diagnostic (host wait spin / sleep), then the bench with the sharded product paths
O=gpurun_out; T=${1:-n8}; N=${2:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
E2E_QUICK=1 timeout 300 $R tools/e2e_ranks.py > $O/e2e_ranks_n${N}_$T.log 2> $O/e2e_ranks_n${N}_$T.err; cat $O/e2e_ranks_n${N}_$T.log; tail -3 $O/e2e_ranks_n${N}_$T.err
timeout 600 $R bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_n${N}_$T.log 2> $O/r02_bench_n${N}_$T.err; tail -c 2500 $O/r02_bench_n${N}_$T.log; tail -3 $O/r02_bench_n${N}_$T.err
RYDCZ_HOST_WAIT=sleep timeout 600 $R bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-configs > $O/r02_bench_n${N}_sleep_$T.log 2> $O/r02_bench_n${N}_sleep_$T.err; tail -c 1200 $O/r02_bench_n${N}_sleep_$T.log
