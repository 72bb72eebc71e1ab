"""One band-scheduled MBConv block (csrc/mbconv_band.cu) and one three-kernel block at the bench shape, for ncu:
    ncu --replay-mode application --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
        -k regex:"mbconv_band|expand_planar|dw_planar|project_planar" python tools/prof_band.py
"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], 'none']
import torch
spec = importlib.util.spec_from_file_location('tp', os.path.join(ROOT, 'tools', 'test_planar.py'))
tp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(tp)
B = tp.B
a = tp.run_block(B.IMPL_PLANAR3, 384, 7, 1, 540, 960, True)
b = tp.run_block(B.IMPL_BAND, 384, 7, 1, 540, 960, True)
torch.cuda.synchronize()
print('identical:', bool(torch.equal(a, b)))
