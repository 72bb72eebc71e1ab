"""Warm per-kernel times of the eager C3 step with SAMPLED sub-networks (torch profiler, CUDA activities only)."""
import sys, runpy, torch
sys.argv = ['bench_train.py', '--steps', '10'] + sys.argv[1:]
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    runpy.run_path('/root/repo/tools/bench_train.py', run_name='__main__')
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=70))
