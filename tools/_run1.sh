set -x
python -m pytest tests -m gpu -x -q -k "bn or train or golden" > gpurun_out/t1_pytest.log 2>&1; echo rc=$? >> gpurun_out/t1_pytest.log
for f in 0 1; do
 OFA_BN_FUSED=$f python tools/bench_train.py --max-subnet --steps 30 > gpurun_out/t1_train_max_f$f.log 2>&1
 OFA_BN_FUSED=$f python tools/bench_train.py --max-subnet --graph --steps 30 > gpurun_out/t1_train_graph_f$f.log 2>&1
 OFA_BN_FUSED=$f python tools/bench_train.py --steps 30 > gpurun_out/t1_train_sampled_f$f.log 2>&1
done
python tools/prof_train.py > gpurun_out/t1_prof_train.log 2>&1
