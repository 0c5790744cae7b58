# A/B of library builds under ab/: tools/gpu_ab.sh <tag> <variant> [<variant> ...]   ("base" = the in-tree library)
O=gpurun_out; T=$1; shift
for v in "$@"; do
  if [ $v = base ]; then unset RYDCZ_LIB_PATH; else export RYDCZ_LIB_PATH=$PWD/ab/librydcz_$v.so; fi
  echo "== $v" >> $O/ab_$T.log
  python tools/cluster_ab.py --modes 1 --tables 2 --iters 10 2>&1 | cut -c1-420 >> $O/ab_$T.log
done
cat $O/ab_$T.log
