O=gpurun_out; T=${1:-y}
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02_tests_all_$T.log 2>&1; tail -6 $O/r02_tests_all_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_sjp8192_$T.csv python tools/prof_sjp.py 8192 > $O/ncu_sjp_$T.log 2>&1; tail -2 $O/ncu_sjp_$T.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 300 $O/r02_bench_$T.log; tail -3 $O/r02_bench_$T.err
