O=gpurun_out; mkdir -p $O; TAG=${1:-x}
python -m pytest tests/test_gpu_cluster.py -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -1 $O/pytest_gpu_$TAG.log
python tools/prof_sjp.py > $O/prof_sjp_plain_$TAG.log 2>&1; tail -1 $O/prof_sjp_plain_$TAG.log
python tools/prof_sjp.py 8192 >> $O/prof_sjp_plain_$TAG.log 2>&1; tail -1 $O/prof_sjp_plain_$TAG.log
