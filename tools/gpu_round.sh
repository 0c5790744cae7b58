#!/bin/bash
# One GPU session: tests, bench, ncu launch list + full captures of the dominant kernels, all SURVEY configs.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu_$TAG.log 2>&1; tail -15 $O/pytest_gpu_$TAG.log
python bench.py > $O/bench_$TAG.log 2>&1 && tail -c 3200 $O/bench_$TAG.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.log 2>&1; tail -c 600 $O/bench_ref_$TAG.log
python tools/cluster_ab.py --modes 0,1 > $O/cluster_ab_$TAG.log 2>&1; cut -c1-420 $O/cluster_ab_$TAG.log
python tools/e2e_breakdown.py > $O/e2e_breakdown_$TAG.log 2>&1; cat $O/e2e_breakdown_$TAG.log
python tools/e2e_sweep.py > $O/e2e_sweep_$TAG.log 2>&1; cat $O/e2e_sweep_$TAG.log
# launch list of the same bench command (per-launch durations; cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launches_$TAG.log 2>&1
# full captures of the dominant kernels: clustered evaluate + coefficient kernel (bench workload), 5-level kernel
ncu --set full --clock-control none --import-source on -k regex:evaluate_cluster_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_clu \
    python tools/prof_eval.py --iters 2 > $O/ncu_clu_$TAG.log 2>&1; tail -1 $O/ncu_clu_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:cluster_coef_kernel -s 1 -c 1 -f -o $O/prof_${TAG}_coef \
    python tools/prof_eval.py --iters 2 > $O/ncu_coef_$TAG.log 2>&1; tail -1 $O/ncu_coef_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:evaluate5_kernel -c 1 -f -o $O/prof_${TAG}_eval5 \
    python tools/prof_eval5.py --iters 1 --points 262144 > $O/ncu_eval5_$TAG.log 2>&1; tail -1 $O/ncu_eval5_$TAG.log
python tools/bench_configs.py > $O/configs_$TAG.jsonl 2> $O/configs_$TAG.err; cat $O/configs_$TAG.jsonl | cut -c1-420
# planned paths against the single-interval path on random batches; launch list of a smooth-JP batch
python tools/stress_paths.py 20000 > $O/stress_paths_$TAG.log 2>&1; tail -1 $O/stress_paths_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_sjp_$TAG.csv \
    python tools/prof_sjp.py > $O/prof_sjp_$TAG.log 2>&1; tail -1 $O/prof_sjp_$TAG.log
