"""One stage of the planar MBConv path at the bench shape, a few launches (for ncu captures).
    python tools/prof_planar.py [expand|dw7|dw5|dw3|project]
"""
import sys
sys.argv = [sys.argv[0]] + (sys.argv[1:] or ['dw7'])
which = sys.argv[1]
import importlib.util, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], 'none']
spec = importlib.util.spec_from_file_location('tp', os.path.join(ROOT, 'tools', 'test_planar.py'))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)
B = tp.B
H, W = 540, 960
P = H * W
call = {'expand': lambda: tp.run_expand(B.OFA_F16, 384, 1, P, timing=True),
        'dw7': lambda: tp.run_dw(B.OFA_F16, 7, 1, 384, H, W, timing=True),
        'dw5': lambda: tp.run_dw(B.OFA_F16, 5, 1, 384, H, W, timing=True),
        'dw3': lambda: tp.run_dw(B.OFA_F16, 3, 1, 192, H, W, timing=True),
        'project': lambda: tp.run_project(B.OFA_F16, 384, 1, P, timing=True)}[which]()
for _ in range(3):
    call()
torch.cuda.synchronize()
print('done', which)
