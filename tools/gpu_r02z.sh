O=gpurun_out; T=${1:-z}
timeout 900 python -m pytest tests -m gpu -q > $O/r02_tests_$T.log 2>&1; tail -4 $O/r02_tests_$T.log

python tools/stress_paths.py > $O/stress_$T.log 2>&1; tail -2 $O/stress_$T.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 200 $O/r02_bench_$T.log; tail -3 $O/r02_bench_$T.err
