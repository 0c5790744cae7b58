# quick GPU check: tests, smoke, a short bench (run on the GPU box)
O=gpurun_out; mkdir -p $O; TAG=${1:-q}
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; tail -1 $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -3
