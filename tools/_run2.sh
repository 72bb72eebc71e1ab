set -x
python tools/bench_train.py --max-subnet --steps 3 --warmup 3 > gpurun_out/t2_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t2_launches_train.csv python tools/bench_train.py --max-subnet --steps 1 --warmup 2 > gpurun_out/t2_ncu1.log 2>&1
for k in wgrad_tc_kernel dw_fast_kernel bn_stats_partial_vec8 bn_bwd_reduce_partial_vec8 bn_bwd_apply_vec8 affine_act_vec8 dw_bwd_filter_rows conv_tc_kernel; do
ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 150 --launch-count 3 -o gpurun_out/t2_$k -f python tools/bench_train.py --max-subnet --steps 1 --warmup 2 > gpurun_out/t2_ncu_$k.log 2>&1
done
python tools/hostprof_train.py > gpurun_out/t2_hostprof.log 2>&1
