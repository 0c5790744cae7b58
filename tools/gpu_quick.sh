O=gpurun_out; mkdir -p $O; TAG=r02f
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -2 $O/pytest_gpu_$TAG.log
python tools/cluster_ab.py --modes 0,1 > $O/cluster_ab_$TAG.log 2>&1; cut -c1-330 $O/cluster_ab_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
