O=gpurun_out; mkdir -p $O; TAG=${1:-x}
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -4 $O/pytest_gpu_$TAG.log
python tools/bench_configs.py 2>/dev/null | grep -i "config3\|config2_bb5_65536_noisy" | cut -c1-420
