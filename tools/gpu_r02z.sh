O=gpurun_out; T=${1:-z}
timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_round2.py tests/test_gpu_api.py -q -x > $O/r02_tests_$T.log 2>&1; tail -4 $O/r02_tests_$T.log
for n in 8192 65536; do for l in 1 4; do echo "== n=$n lanes=$l" >> $O/sjp_lanes_$T.log; RYDCZ_REPEAT_LANES=$l python tools/prof_sjp.py $n >> $O/sjp_lanes_$T.log 2>&1; done; done; cat $O/sjp_lanes_$T.log
python tools/stress_paths.py > $O/stress_$T.log 2>&1; tail -4 $O/stress_$T.log
