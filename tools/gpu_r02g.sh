O=gpurun_out; T=${1:-g}
python -m pytest tests/test_gpu_table.py tests/test_gpu_round2.py tests/test_gpu_cluster.py -x -q > $O/r02_tests_$T.log 2>&1; tail -5 $O/r02_tests_$T.log
python tools/cluster_ab.py --modes 1 --tables 0,1,2 > $O/cluster_ab_r02$T.log 2>&1; cut -c1-600 $O/cluster_ab_r02$T.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 1500 $O/r02_bench_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $O/ncu_$T.log 2>&1; tail -2 $O/ncu_$T.log
