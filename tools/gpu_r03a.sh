# round 2, session 4: closed-form single-atom rho (rydcz_closed.cuh) -- GPU suite, A/B against the Chebyshev build, bench
O=gpurun_out; T=${1:-r03a}
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02_tests_$T.log 2>&1; tail -4 $O/r02_tests_$T.log
bash tools/gpu_ab.sh $T base norho01 base
python bench.py > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 300 $O/r02_bench_$T.log; tail -3 $O/r02_bench_$T.err
