O=gpurun_out; T=${1:-j}
python -m pytest tests -m gpu -x -q > $O/r02_tests_$T.log 2>&1; tail -4 $O/r02_tests_$T.log
python bench.py --steps 12 --warmup 3 --no-cpu-baseline > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 600 $O/r02_bench_$T.log
ncu --set full --clock-control none --import-source on -k regex:prologue_kernel -s 1 -c 1 -f -o $O/r02_prologue_$T python tools/prof_eval.py --iters 2 > $O/ncu_pro_$T.log 2>&1; tail -1 $O/ncu_pro_$T.log
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 1 -c 1 -f -o $O/r02_plan_$T python tools/prof_eval.py --iters 2 > $O/ncu_plan_$T.log 2>&1; tail -1 $O/ncu_plan_$T.log
ncu --set full --clock-control none --import-source on -k regex:cluster_table_kernel -s 1 -c 1 -f -o $O/r02_table_$T python tools/prof_eval.py --iters 2 > $O/ncu_tab_$T.log 2>&1; tail -1 $O/ncu_tab_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 64 --csv --log-file $O/r02_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > $O/ncu_$T.log 2>&1; tail -c 200 $O/ncu_$T.log
