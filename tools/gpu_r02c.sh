O=gpurun_out; T=${1:-c}
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 1 -c 1 -f -o $O/r02_eval_cluster_general_$T python tools/prof_shots.py > $O/ncu_shots_$T.log 2>&1; tail -2 $O/ncu_shots_$T.log
ncu --set full --clock-control none --import-source on -k regex:repeat_apply4_kernel -s 1 -c 1 -f -o $O/r02_repeat_apply4_$T python tools/prof_sjp.py 65536 > $O/ncu_sjp4_$T.log 2>&1; tail -2 $O/ncu_sjp4_$T.log
