# ncu --set full of the step's three largest kernels (planner, prologue, clustered evaluate) on the bench workload
O=gpurun_out; T=${1:-r03b}
python tools/prof_eval.py --iters 4 > $O/prof_plain_$T.log 2>&1; tail -1 $O/prof_plain_$T.log
ncu --set full --clock-control none --import-source on -k 'regex:^(plan_kernel|prologue_kernel|evaluate_cluster_kernel)$' -s 6 -c 3 -f -o $O/r02_step_kernels_$T python tools/prof_eval.py --iters 4 > $O/ncu_step_$T.log 2>&1; tail -2 $O/ncu_step_$T.log
