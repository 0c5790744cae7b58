O=gpurun_out; T=${1:-v}
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_tests_all_$T.log 2>&1; tail -6 $O/r02_tests_all_$T.log
python bench.py > $O/r02_bench_full_$T.log 2> $O/r02_bench_full_$T.err; tail -c 600 $O/r02_bench_full_$T.log; tail -3 $O/r02_bench_full_$T.err
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 3 -c 1 -f -o $O/r02_eval_cluster_$T python tools/prof_eval.py --iters 2 > $O/ncu_clu_$T.log 2>&1; tail -1 $O/ncu_clu_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/ncu_$T.log 2>&1; tail -c 300 $O/ncu_$T.log
python tools/e2e_sweep.py > $O/e2e_sweep_$T.log 2>&1; cat $O/e2e_sweep_$T.log
