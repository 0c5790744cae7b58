O=gpurun_out; T=${1:-l}
python -m pytest tests/test_gpu_table.py tests/test_gpu_round2.py tests/test_gpu_cluster.py tests/test_gpu_api.py tests/test_gpu_parity.py -x -q > $O/r02_tests_$T.log 2>&1; tail -3 $O/r02_tests_$T.log
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 300 $O/r02_bench_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file $O/r02_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/ncu_$T.log 2>&1; tail -c 100 $O/ncu_$T.log
python tools/e2e_breakdown.py > $O/e2e_breakdown_r02$T.log 2>&1; cat $O/e2e_breakdown_r02$T.log
python tools/e2e_sweep.py > $O/e2e_sweep_r02$T.log 2>&1; cat $O/e2e_sweep_r02$T.log
