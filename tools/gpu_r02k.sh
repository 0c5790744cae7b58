O=gpurun_out; T=${1:-k}
python -m pytest tests/test_gpu_table.py tests/test_gpu_round2.py tests/test_gpu_cluster.py tests/test_gpu_api.py -x -q > $O/r02_tests_$T.log 2>&1; tail -3 $O/r02_tests_$T.log
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 700 $O/r02_bench_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file $O/r02_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/ncu_$T.log 2>&1; tail -c 200 $O/ncu_$T.log
