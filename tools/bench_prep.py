"""Device-side SR batch preparation (ofa_b200.data.SRTrainBatchPrep) at the C3 shape -- 64 patches of 96x96 cut from
64 uint8 source images -- timed with CUDA events, beside the same chain on the host with Pillow / torchvision (what
the reference's data-loader workers run, div2k_setxx.py:166-171,288-298), one process, one thread.
    python tools/bench_prep.py [--n 64] [--src 339x510] [--size 96]"""
import argparse, os, sys, time
import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'ofa-for-super-resolution_b200'))
from ofa_b200 import data as D, backend as B

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=64)
ap.add_argument('--src', default='678x1020')
ap.add_argument('--size', type=int, default=96)
a = ap.parse_args()
H, W = [int(v) for v in a.src.split('x')]
dev = torch.device('cuda:0')
rs = np.random.RandomState(0)
src_np = rs.randint(0, 256, (a.n, H, W, 3), dtype=np.uint8)
src = torch.from_numpy(src_np).to(dev)
torch.manual_seed(0)
params = D.sample_train_params(a.n, H, W, a.size)
prep = D.SRTrainBatchPrep(a.size)
for _ in range(3):
    out = prep(src, params)
torch.cuda.synchronize()
B.lib().ofa_launch_count_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 20
for _ in range(iters):
    out = prep(src, params)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
launches = B.lib().ofa_launch_count() // iters
out_bytes = sum(v.numel() * 4 for v in out.values())
print(f'GPU  prep: {ms * 1e3:8.1f} us / batch of {a.n} ({a.n / ms * 1e3:,.0f} samples/s, {launches} launches, '
      f'{out_bytes / 1e6:.1f} MB of fp32 tensors out)')
try:
    from PIL import Image
    import torchvision.transforms as T
    t0 = time.perf_counter()
    for n in range(a.n):
        i, j, flip, ang = params[n]
        p = Image.fromarray(src_np[n], 'RGB').crop((j, i, j + a.size, i + a.size))
        if flip:
            p = p.transpose(Image.FLIP_LEFT_RIGHT)
        p = p.rotate(ang)
        l2 = p.resize((a.size // 2, a.size // 2), Image.BICUBIC)
        l4 = p.resize((a.size // 4, a.size // 4), Image.BICUBIC)
        ref = [T.ToTensor()(q) for q in (p, l2, l4)]
    dt = time.perf_counter() - t0
    ok = torch.equal(ref[0], out['image'][a.n - 1].cpu()) and torch.equal(ref[2], out['4x_down_image'][a.n - 1].cpu())
    print(f'CPU  prep (Pillow + torchvision, 1 thread): {dt * 1e3:8.2f} ms / batch ({a.n / dt:,.0f} samples/s); '
          f'last sample bit-equal to the GPU result: {ok}')
except ImportError as e:
    print('Pillow / torchvision not importable:', e)
