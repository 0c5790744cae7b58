O=gpurun_out; T=${1:-p}
timeout 900 python -m pytest tests/test_gpu_five_level.py -x -q > $O/r02_tests5_$T.log 2>&1; tail -15 $O/r02_tests5_$T.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_tests_all_$T.log 2>&1; tail -5 $O/r02_tests_all_$T.log
