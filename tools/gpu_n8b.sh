O=gpurun_out; T=${1:-n8}; N=${2:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
E2E_QUICK=1 timeout 300 $R tools/e2e_ranks.py > $O/e2e_ranks_n${N}_$T.log 2> $O/e2e_ranks_n${N}_$T.err; cat $O/e2e_ranks_n${N}_$T.log; tail -2 $O/e2e_ranks_n${N}_$T.err
