O=gpurun_out; mkdir -p $O; TAG=${1:-x}
python -m pytest tests/test_gpu_api.py tests/test_gpu_cluster.py -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -3 $O/pytest_gpu_$TAG.log
python tools/e2e_breakdown.py > $O/e2e_breakdown_$TAG.log 2>&1; cat $O/e2e_breakdown_$TAG.log
python tools/e2e_sweep.py > $O/e2e_sweep_$TAG.log 2>&1; cat $O/e2e_sweep_$TAG.log
