O=gpurun_out; T=${1:-b}
nvidia-smi --query-gpu=name,temperature.gpu,clocks.sm --format=csv > $O/smi_$T.log 2>&1; cat $O/smi_$T.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 200 $O/r02_bench_$T.log; tail -3 $O/r02_bench_$T.err
timeout 600 python -m pytest tests/test_gpu_round2.py -q -x > $O/r02_tests_$T.log 2>&1; tail -3 $O/r02_tests_$T.log
python tools/stress_paths.py > $O/stress_$T.log 2>&1; tail -2 $O/stress_$T.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity --no-configs > $O/r02_bench2_$T.log 2> $O/r02_bench2_$T.err; tail -c 200 $O/r02_bench2_$T.log; tail -3 $O/r02_bench2_$T.err
