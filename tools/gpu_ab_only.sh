O=gpurun_out; mkdir -p $O; TAG=${1:-x}
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -5 $O/pytest_gpu_$TAG.log
python tools/cluster_ab.py --modes 1 > $O/cluster_ab_$TAG.log 2>&1; cut -c1-330 $O/cluster_ab_$TAG.log
python tools/bench_configs.py 2>/dev/null | grep -i "config4\|config1_lp_65536_distinct" | cut -c1-400
