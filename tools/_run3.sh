python tools/hosttime_train.py > gpurun_out/t3_host_sampled.log 2>&1
python tools/hosttime_train.py --max-subnet > gpurun_out/t3_host_max.log 2>&1
for k in wgrad_tc_kernel dw_fast_kernel bn_stats_partial_vec8 bn_bwd_reduce_partial_vec8 bn_bwd_apply_vec8 affine_act_vec8 dw_bwd_filter_rows; do
ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 30 --launch-count 3 -o gpurun_out/t2_$k -f python tools/bench_train.py --max-subnet --steps 1 --warmup 2 > gpurun_out/t2_ncu_$k.log 2>&1
done
