"""Golden vectors for the SR data preparation: runs Pillow + torchvision themselves, through the
reference's own transform helpers when /root/reference is importable, on seeded uint8 images and stores the outputs.
    python tests/golden/make_golden_prep.py        (build container only; writes reference_prep.npz)
"""
import os, sys
import numpy as np
from PIL import Image
import torchvision.transforms as T

sys.path.insert(0, '/root/reference')
from ofa.imagenet_codebase.data_providers.div2k_setxx import get_transform_L, crop   # the reference's Scale / crop

rs = np.random.RandomState(0)
out = {}
# smooth + noisy content so that the cubic lobes clip on both sides
def synth(h, w, seed):
    r = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(xx / 7.0 + seed) * np.cos(yy / 5.0) + 60 * r.randn(h, w)
    img = np.stack([base, base[::-1] * 0.7 + 40 * r.randn(h, w), 255 - base + 30 * r.randn(h, w)], -1)
    return np.clip(img, 0, 255).astype(np.uint8)

cases = [(96, 96, 1), (64, 48, 2), (50, 70, 3), (33, 21, 4), (8, 8, 5), (120, 100, 6)]
for h, w, seed in cases:
    img = synth(h, w, seed)
    pil = Image.fromarray(img, 'RGB')
    out['img_%dx%d' % (h, w)] = img
    for opt in (2, 4, 8):        # get_transform_L accepts 2, 4 and 8 (div2k_setxx.py:382); the data sets use 2 and 4
        lo = get_transform_L(opt=opt)(pil)
        out['down%d_%dx%d' % (opt, h, w)] = np.asarray(lo)
        out['down%d_tensor_%dx%d' % (opt, h, w)] = T.ToTensor()(lo).numpy()
    out['tensor_%dx%d' % (h, w)] = T.ToTensor()(pil).numpy()

# crop / flip / rotate chain as train_transforms applies it (div2k_setxx.py:166-171), explicit parameters
img = synth(120, 100, 6)
pil = Image.fromarray(img, 'RGB')
angles = [0.0, 90.0, -90.0, 37.5, -12.25, 89.999, -64.0, 45.0, 1e-3, 180.0, -33.3, 71.7]
params = []
for k, ang in enumerate(angles):
    i, j, flip = int(rs.randint(0, 120 - 48 + 1)), int(rs.randint(0, 100 - 48 + 1)), int(rs.randint(0, 2))
    p = crop(pil, i, j, 48, 48)
    if flip:
        p = p.transpose(Image.FLIP_LEFT_RIGHT)
    p = p.rotate(ang)                       # RandomRotation -> F.rotate(img, angle) with NEAREST, expand False, fill 0
    out['aug_%d' % k] = np.asarray(p)
    params.append((i, j, flip, ang))
# non-square rotation
for k, ang in enumerate([30.0, -75.5, 90.0]):
    out['rot_rect_%d' % k] = np.asarray(Image.fromarray(synth(40, 64, 9), 'RGB').rotate(ang))
out['rot_rect_img'] = synth(40, 64, 9)
out['rot_rect_angles'] = np.array([30.0, -75.5, 90.0])
out['aug_params'] = np.array(params, np.float64)
import PIL, torchvision
out['versions'] = np.array(['Pillow ' + PIL.__version__, 'torchvision ' + torchvision.__version__])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_prep.npz'), **out)
print('wrote', len(out), 'arrays')
