O=gpurun_out; T=${1:-n2}; N=${2:-2}
python -m pytest tests/test_gpu_sharded_nccl.py -x -q > $O/r02_tests_$T.log 2>&1; tail -3 $O/r02_tests_$T.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 1500 $O/r02_bench_$T.log; tail -5 $O/r02_bench_$T.err
