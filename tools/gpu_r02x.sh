O=gpurun_out; T=${1:-x}
for n in 256 8192; do echo "== $n" >> $O/small_$T.log; python tools/e2e_breakdown.py $n >> $O/small_$T.log 2>&1; done; cat $O/small_$T.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_sjp8192_$T.csv python tools/prof_sjp.py 8192 > $O/ncu_sjp_$T.log 2>&1; tail -2 $O/ncu_sjp_$T.log
