O=gpurun_out; T=${1:-n}
python -m pytest tests/test_gpu_api.py tests/test_gpu_round2.py tests/test_gpu_table.py -x -q > $O/r02_tests_$T.log 2>&1; tail -3 $O/r02_tests_$T.log
python tools/e2e_ranks.py > $O/e2e_ranks_n1_$T.log 2>&1; cat $O/e2e_ranks_n1_$T.log
E2E_PINNED=0 python tools/e2e_ranks.py > $O/e2e_ranks_n1_pageable_$T.log 2>&1; head -3 $O/e2e_ranks_n1_pageable_$T.log
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-configs > $O/r02_bench_$T.log 2> $O/r02_bench_$T.err; tail -c 300 $O/r02_bench_$T.log; tail -3 $O/r02_bench_$T.err
