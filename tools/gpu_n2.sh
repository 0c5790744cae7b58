O=gpurun_out; T=${1:-n2}
timeout 600 python -m pytest tests/test_gpu_sharded_nccl.py -q > $O/r02_tests_nccl_$T.log 2>&1; tail -4 $O/r02_tests_nccl_$T.log
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $R bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_n2_$T.log 2> $O/r02_bench_n2_$T.err; tail -c 400 $O/r02_bench_n2_$T.log; tail -3 $O/r02_bench_n2_$T.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r02_bench_ref_$T.log 2>&1; tail -c 600 $O/r02_bench_ref_$T.log
