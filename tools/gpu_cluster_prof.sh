# clustered path: A/B timing, then full ncu captures of its two kernels (run on the GPU box)
TAG=${1:-r01d}
O=gpurun_out
mkdir -p $O
python tools/cluster_ab.py --modes 0,1 > $O/cluster_ab_$TAG.log 2>&1; cat $O/cluster_ab_$TAG.log | cut -c1-500
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_clu \
    python tools/prof_eval.py --iters 2 > $O/ncu_clu_$TAG.log 2>&1; tail -2 $O/ncu_clu_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:cluster_coef_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_coef \
    python tools/prof_eval.py --iters 2 > $O/ncu_coef_$TAG.log 2>&1; tail -2 $O/ncu_coef_$TAG.log
