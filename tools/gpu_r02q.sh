O=gpurun_out; T=${1:-q}
timeout 900 python -m pytest tests/test_gpu_five_level.py -x -q > $O/r02_tests5_$T.log 2>&1; tail -5 $O/r02_tests5_$T.log
for v in base cl4 pl4 pl2; do
  if [ $v = base ]; then unset RYDCZ_LIB_PATH; else export RYDCZ_LIB_PATH=$PWD/ab/librydcz_$v.so; fi
  echo "== $v" >> $O/ab_$T.log
  python tools/cluster_ab.py --modes 1 --tables 1,2 --iters 10 2>&1 | cut -c1-420 >> $O/ab_$T.log
done
cat $O/ab_$T.log
