python -m pytest tests -m gpu -x -q -k "block_training or train or golden or adam" > gpurun_out/t4_pytest.log 2>&1; echo rc=$? >> gpurun_out/t4_pytest.log
python tools/hosttime_train.py > gpurun_out/t4_host_sampled.log 2>&1
python tools/hosttime_train.py --max-subnet > gpurun_out/t4_host_max.log 2>&1
python tools/bench_train.py --max-subnet --graph --steps 30 > gpurun_out/t4_train_graph.log 2>&1
OFA_BLOCK_TRAIN=0 python tools/bench_train.py --max-subnet --graph --steps 30 > gpurun_out/t4_train_graph_layerwise.log 2>&1
python tools/bench_train.py --max-subnet --graph --steps 30 > gpurun_out/t4_train_graph2.log 2>&1
OFA_BLOCK_TRAIN=0 python tools/bench_train.py --max-subnet --graph --steps 30 > gpurun_out/t4_train_graph_layerwise2.log 2>&1
