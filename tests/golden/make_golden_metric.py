"""Golden values of the reference's evaluation metric — psnr(rgb2y(tensor2img_np(a)), rgb2y(tensor2img_np(b))) —
sr_run_manager.py:364,496,567-597 + ofa/utils.py:27-34, produced by importing the UNMODIFIED reference functions
(CPU).  Run once in the build container:   python tests/golden/make_golden_metric.py
Inputs come from numpy RandomState seeds, so only seeds / shapes / PSNR values are stored."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
from ofa.imagenet_codebase.run_manager.sr_run_manager import tensor2img_np, rgb2y  # noqa: E402
from ofa.utils import psnr  # noqa: E402

cases = []
for seed, shape, noise in [(1, (1, 3, 37, 53), 0.05), (2, (3, 3, 24, 40), 0.02), (3, (4, 3, 16, 16), 0.2), (4, (7, 3, 9, 21), 0.01),
                           (5, (1, 3, 64, 64), 0.0)]:
    rs = np.random.RandomState(seed)
    a = (rs.rand(*shape) * 1.2 - 0.1).astype(np.float32)          # includes values outside [0, 1] (clamped)
    b = (a + noise * rs.randn(*shape)).astype(np.float32)
    v = psnr(rgb2y(tensor2img_np(torch.from_numpy(a))), rgb2y(tensor2img_np(torch.from_numpy(b))))
    cases.append({'seed': seed, 'shape': list(shape), 'noise': noise, 'psnr': (None if v == float('inf') else v)})
with open(os.path.join(HERE, 'reference_metric.json'), 'w') as f:
    json.dump(cases, f, indent=1)
print(cases)
