O=gpurun_out; T=${1:-d}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:evaluate_cluster_kernel<\(int\)3, \(bool\)1, \(bool\)0>' -s 1 -c 1 -f -o $O/r02_eval_cluster_general_$T python tools/prof_shots.py > $O/ncu_shots_$T.log 2>&1; tail -3 $O/ncu_shots_$T.log | cut -c1-200
