# clustered path: GPU tests, A/B timing, optional ncu captures (run on the GPU box)
TAG=${1:-r01f}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu_$TAG.log 2>&1; tail -6 $O/pytest_gpu_$TAG.log
python tools/cluster_ab.py --modes 0,1 > $O/cluster_ab_$TAG.log 2>&1; cat $O/cluster_ab_$TAG.log | cut -c1-420
python tools/cluster_ab.py --modes 0,1 --points 65536 >> $O/cluster_ab_$TAG.log 2>&1; tail -2 $O/cluster_ab_$TAG.log | cut -c1-420
python tools/cluster_ab.py --modes 0,1 --noiseless >> $O/cluster_ab_$TAG.log 2>&1; tail -2 $O/cluster_ab_$TAG.log | cut -c1-420
if [ "$2" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_clu \
    python tools/prof_eval.py --iters 2 > $O/ncu_clu_$TAG.log 2>&1; tail -1 $O/ncu_clu_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:cluster_coef_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_coef \
    python tools/prof_eval.py --iters 2 > $O/ncu_coef_$TAG.log 2>&1; tail -1 $O/ncu_coef_$TAG.log
fi
