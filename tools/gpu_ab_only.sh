O=gpurun_out; mkdir -p $O; TAG=${1:-x}
python tools/cluster_ab.py --modes 1 > $O/cluster_ab_$TAG.log 2>&1; cut -c1-330 $O/cluster_ab_$TAG.log
python tools/cluster_ab.py --modes 1 --points 65536 >> $O/cluster_ab_$TAG.log 2>&1; tail -1 $O/cluster_ab_$TAG.log | cut -c1-330
