O=gpurun_out; T=${1:-s}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_tests_all_$T.log 2>&1; tail -8 $O/r02_tests_all_$T.log
bash tools/gpu_ab.sh $T base noclosed
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-configs > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 1500 $O/r02_bench_$T.log; tail -3 $O/r02_bench_$T.err
