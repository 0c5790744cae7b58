O=gpurun_out; mkdir -p $O; TAG=${1:-x}
python tools/cluster_ab.py --modes 1 > $O/cluster_ab_$TAG.log 2>&1; cut -c1-330 $O/cluster_ab_$TAG.log
python tools/cluster_ab.py --modes 1 --points 65536 >> $O/cluster_ab_$TAG.log 2>&1; tail -1 $O/cluster_ab_$TAG.log | cut -c1-330
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_clu \
    python tools/prof_eval.py --iters 2 > $O/ncu_clu_$TAG.log 2>&1; tail -1 $O/ncu_clu_$TAG.log
