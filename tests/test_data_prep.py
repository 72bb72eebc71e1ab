"""SR data preparation: the numpy oracle against Pillow's own outputs (reference_prep.npz, made by
tests/golden/make_golden_prep.py through the reference's Scale / crop helpers), the host-side logic of ofa_b200.data on
the CPU, and (-m gpu) the CUDA kernels against both.  Everything here is byte / integer work: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

import sr_data_prep_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ['96x96', '64x48', '50x70', '33x21', '8x8', '120x100']


@pytest.fixture(scope='module')
def prep():
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_prep.npz'))


# ------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize('hw', CASES)
def test_oracle_bicubic_matches_pillow(prep, hw):
    img = prep['img_' + hw]
    for opt in (2, 4, 8):
        got = P.scale_down(img, opt)
        assert np.array_equal(got, prep['down%d_%s' % (opt, hw)])
        assert np.array_equal(P.to_tensor(got), prep['down%d_tensor_%s' % (opt, hw)])
    assert np.array_equal(P.to_tensor(img), prep['tensor_' + hw])


def test_oracle_augment_matches_pillow(prep):
    img = prep['img_120x100']
    for k, (i, j, flip, ang) in enumerate(prep['aug_params']):
        h = P.crop(img, int(i), int(j), 48, 48)
        if flip:
            h = P.hflip(h)
        assert np.array_equal(P.rotate_nearest(np.ascontiguousarray(h), float(ang)), prep['aug_%d' % k]), (k, ang)
    for k, ang in enumerate(prep['rot_rect_angles']):
        assert np.array_equal(P.rotate_nearest(prep['rot_rect_img'], float(ang)), prep['rot_rect_%d' % k])


@pytest.mark.parametrize('sizes', [(96, 48), (96, 24), (50, 25), (50, 12), (33, 8), (21, 5), (8, 2), (7, 7), (17, 40),
                                   (2040, 510), (1, 1)])
def test_host_resample_table_matches_oracle(sizes):
    """ofa_resample_build_table (C++, double precision on the host) == Pillow's precompute_coeffs as restated."""
    from ofa_b200 import data as D
    t = D._ResampleTable(sizes[0], sizes[1], None)
    ks, bounds, kk = P.resample_coeffs(*sizes)
    assert ks == t.ksize
    assert np.array_equal(bounds, t.bounds_host.numpy()) and np.array_equal(kk, t.kk_host.numpy())


def test_rotation_params_modes():
    from ofa_b200 import data as D
    assert D.rotation_params(0.0, 48)[0] == 0 and D.rotation_params(360.0, 48)[0] == 0
    assert D.rotation_params(180.0, 48)[0] == 1 and D.rotation_params(-180.0, 48)[0] == 1
    assert D.rotation_params(90.0, 48)[0] == 2 and D.rotation_params(-270.0, 48)[0] == 2
    assert D.rotation_params(-90.0, 48)[0] == 3 and D.rotation_params(270.0, 48)[0] == 3
    mode, a0, a1, a2, a3, a4, a5 = D.rotation_params(45.0, 48)
    assert mode == 4 and a0 == a4 and a1 == -a3 and a0 == 46341


def test_sampled_params_follow_torchvision_rng_order():
    """sample_train_params draws what RandomCrop / RandomHorizontalFlip / RandomRotation would draw from the same
    torch seed (div2k_setxx.py:166-171 composes exactly these three)."""
    tv = pytest.importorskip('torchvision.transforms')
    from PIL import Image
    from ofa_b200 import data as D
    img = np.random.RandomState(3).randint(0, 256, (60, 70, 3), dtype=np.uint8)
    torch.manual_seed(11)
    params = D.sample_train_params(5, 60, 70, 32)
    torch.manual_seed(11)
    chain = tv.Compose([tv.RandomCrop(32), tv.RandomHorizontalFlip(), tv.RandomRotation(degrees=(-90, 90))])
    for (i, j, flip, ang) in params:
        ref = np.asarray(chain(Image.fromarray(img, 'RGB')))
        h = P.crop(img, i, j, 32, 32)
        if flip:
            h = P.hflip(h)
        assert np.array_equal(P.rotate_nearest(np.ascontiguousarray(h), ang), ref)


# ------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


@pytest.mark.gpu
@pytest.mark.parametrize('hw', CASES)
def test_gpu_bicubic_matches_pillow(dev, prep, hw):
    from ofa_b200 import data as D
    img = torch.from_numpy(prep['img_' + hw]).to(dev)[None].contiguous()
    for opt in (2, 4, 8):
        ref = prep['down%d_%s' % (opt, hw)]
        got, got_u8 = D.bicubic_resize(img, ref.shape[0], ref.shape[1], want_u8=True)
        assert np.array_equal(got_u8[0].cpu().numpy(), ref)
        assert np.array_equal(got[0].cpu().numpy(), prep['down%d_tensor_%s' % (opt, hw)])     # ToTensor, bit-exact
        assert np.array_equal(D.scale_down(img, opt)[0].cpu().numpy(), prep['down%d_tensor_%s' % (opt, hw)])


@pytest.mark.gpu
def test_gpu_augment_matches_pillow(dev, prep):
    from ofa_b200 import data as D
    params = [(int(i), int(j), bool(f), float(a)) for (i, j, f, a) in prep['aug_params']]
    src = torch.from_numpy(prep['img_120x100']).to(dev)[None].repeat(len(params), 1, 1, 1).contiguous()
    hr, hr_u8 = D.augment(src, params, 48)
    for k in range(len(params)):
        assert np.array_equal(hr_u8[k].cpu().numpy(), prep['aug_%d' % k]), (k, params[k])
        assert np.array_equal(hr[k].cpu().numpy(), P.to_tensor(prep['aug_%d' % k]))


@pytest.mark.gpu
def test_gpu_train_batch_prep_vs_oracle_full_size(dev):
    """BASELINE.json configs[2] shape: 64 patches of 96x96 (+ 48x48 and 24x24 LR) from 64 different source images,
    random crop / flip / rotation parameters; every output tensor bit-equal to the oracle's."""
    from ofa_b200 import data as D
    rs = np.random.RandomState(5)
    src = rs.randint(0, 256, (64, 130, 150, 3), dtype=np.uint8)
    torch.manual_seed(3)
    params = D.sample_train_params(64, 130, 150, 96)
    params[0] = (0, 0, False, 0.0)
    params[1] = (34, 54, True, 90.0)
    params[2] = (34, 54, True, -90.0)
    batch = D.SRTrainBatchPrep(96)(torch.from_numpy(src).to(dev), params)
    assert batch['image'].shape == (64, 3, 96, 96) and batch['4x_down_image'].shape == (64, 3, 24, 24)
    for n in (0, 1, 2, 3, 17, 40, 63):
        ref = P.prepare_sample(src[n], params[n][0], params[n][1], 96, params[n][2], params[n][3])
        for key in ('image', '2x_down_image', '4x_down_image'):
            assert np.array_equal(batch[key][n].cpu().numpy(), ref[key]), (n, key)


@pytest.mark.gpu
def test_gpu_resize_properties_large(dev):
    """Size-independent properties at a validation-frame size (DIV2K 2040x1356): a constant image stays constant (the
    window weights sum to 2^22 after Pillow's normalisation up to rounding: allow +-1), a horizontally flipped image
    resizes to the flipped result (the coefficient windows are mirror images), and identity size is a copy."""
    from ofa_b200 import data as D
    const = torch.full((1, 1356, 2040, 3), 131, dtype=torch.uint8, device=dev)
    _, c8 = D.bicubic_resize(const, 339, 510, want_u8=True)
    assert int((c8.int() - 131).abs().max()) <= 1
    rs = np.random.RandomState(9)
    img = torch.from_numpy(rs.randint(0, 256, (1, 1356, 2040, 3), dtype=np.uint8)).to(dev)
    _, a = D.bicubic_resize(img, 678, 1020, want_u8=True)
    _, b = D.bicubic_resize(img.flip(2).contiguous(), 678, 1020, want_u8=True)
    assert torch.equal(a.flip(2), b)
    _, same = D.bicubic_resize(img, 1356, 2040, want_u8=True)
    assert torch.equal(same, img)
    # a row band of the big frame against the oracle (the oracle is pure python over rows: keep it small)
    band = img[:, :64, :256].contiguous()
    _, got = D.bicubic_resize(band, 16, 64, want_u8=True)
    assert np.array_equal(got[0].cpu().numpy(), P.bicubic_resize_u8(band[0].cpu().numpy(), 16, 64))


@pytest.mark.gpu
def test_gpu_prep_bad_arguments(dev):
    from ofa_b200 import data as D
    with pytest.raises(RuntimeError):
        D.bicubic_resize(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), 4, 4)          # CPU tensor: no CPU path
    src = torch.zeros(1, 20, 20, 3, dtype=torch.uint8, device=dev)
    with pytest.raises(ValueError):
        D.augment(src, [(10, 0, False, 0.0)], 16)                                     # crop leaves the image
    out = D.augment(src[:0], [], 16)[0]
    assert out.shape == (0, 3, 16, 16)
