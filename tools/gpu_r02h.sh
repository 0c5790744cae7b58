O=gpurun_out; T=${1:-h}
python -m pytest tests/test_gpu_table.py tests/test_gpu_round2.py tests/test_gpu_cluster.py -x -q > $O/r02_tests_$T.log 2>&1; tail -5 $O/r02_tests_$T.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 1200 $O/r02_bench_$T.log
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 1 -c 1 -f -o $O/r02_plan_$T python tools/prof_eval.py --iters 2 > $O/ncu_plan_$T.log 2>&1; tail -1 $O/ncu_plan_$T.log
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 3 -c 1 -f -o $O/r02_eval_cluster_$T python tools/prof_eval.py --iters 2 > $O/ncu_clu_$T.log 2>&1; tail -1 $O/ncu_clu_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/ncu_$T.log 2>&1; tail -c 300 $O/ncu_$T.log
