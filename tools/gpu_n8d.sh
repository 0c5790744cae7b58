# eight GPUs, final build of round 2: the bench line (weak-scaling headline + sharded product paths), then the NCCL test
O=gpurun_out; T=${1:-n8d}; N=${2:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $R bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_n${N}_$T.log 2> $O/r02_bench_n${N}_$T.err; tail -c 300 $O/r02_bench_n${N}_$T.log; tail -3 $O/r02_bench_n${N}_$T.err
timeout 200 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q > $O/r02_tests_nccl_$T.log 2>&1; tail -3 $O/r02_tests_nccl_$T.log
