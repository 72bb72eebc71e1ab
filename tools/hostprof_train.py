"""Host-side (Python) profile of the eager training step: cProfile over a few steps of tools/bench_train.py."""
import cProfile, pstats, sys, runpy, io
sys.argv = ['bench_train.py', '--steps', '20', '--warmup', '4']
pr = cProfile.Profile()
pr.enable()
runpy.run_path('/root/repo/tools/bench_train.py', run_name='__main__')
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(45)
print(s.getvalue()[:9000])
