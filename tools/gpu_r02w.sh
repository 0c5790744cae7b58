O=gpurun_out; T=${1:-w}
timeout 1500 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_round2.py tests/test_gpu_table.py tests/test_gpu_five_level.py -q > $O/r02_tests_$T.log 2>&1; tail -6 $O/r02_tests_$T.log
python tools/overlap_probe.py > $O/overlap_$T.log 2>&1; cat $O/overlap_$T.log
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 2 -c 1 -f -o $O/r02_eval_cluster_$T python tools/prof_eval.py --iters 4 > $O/ncu_clu_$T.log 2>&1; tail -2 $O/ncu_clu_$T.log
