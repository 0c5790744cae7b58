# final verification of a build: the whole GPU suite, smoke(), the default bench line, the reference arm
O=gpurun_out; T=${1:-final}
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_tests_all_$T.log 2>&1; tail -3 $O/r02_tests_all_$T.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; tail -2 $O/smoke_$T.log
python bench.py > $O/r02_bench_default_$T.log 2> $O/r02_bench_default_$T.err; tail -c 300 $O/r02_bench_default_$T.log; tail -3 $O/r02_bench_default_$T.err
