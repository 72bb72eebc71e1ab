"""`ofa` import shim (ofa_b200/compat.py) -- CPU: the reference's import lines resolve to the drop-in classes."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'ofa-for-super-resolution_b200')


def _run(code):
    return subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300,
                          env=dict(os.environ, PYTHONPATH=PKG))


def test_standalone_overlay_resolves_the_reference_import_lines():
    r = _run('''
import ofa_b200.compat as c
c.install_ofa_overlay()
from ofa.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d, DynamicPointConv2d, DynamicBatchNorm2d
from ofa.elastic_nn.modules.dynamic_layers import DynamicMBConvLayer
from ofa.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
from ofa.layers import ConvLayer, MBInvertedConvLayer, IdentityLayer
from ofa.utils import make_divisible, int2list
from ofa.elastic_nn.utils import set_running_statistics
from ofa.elastic_nn.training.progressive_shrinking import train_one_epoch
assert all(k.__module__.startswith('ofa_b200.') for k in (DynamicSeparableConv2d, DynamicMBConvLayer, OFAMobileNetS4, ConvLayer))
net = OFAMobileNetS4(ks_list=[3, 5, 7], expand_ratio_list=[3, 4, 6], depth_list=[2, 3, 4], pixelshuffle_depth_list=[1, 2])
net.set_active_subnet(ks=7, e=6, d=4, pixel_d=2)
print('ok', len(net.state_dict()))
''')
    assert r.returncode == 0 and r.stdout.startswith('ok'), r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir('/root/reference/ofa'), reason='no reference checkout on this machine')
def test_overlay_keeps_the_reference_for_everything_outside_the_hot_path():
    r = _run('''
import ofa_b200.compat as c
c.install_ofa_overlay('/root/reference')
from ofa.elastic_nn.networks import OFAMobileNetX4
from ofa.elastic_nn.modules.dynamic_op import DynamicSeparableConv2d
from ofa.elastic_nn.training.progressive_shrinking import train_one_epoch, load_models, validate
from ofa.imagenet_codebase.run_manager.sr_run_manager import SRRunManager, RunConfig
from ofa.utils import download_url, AverageMeter
assert OFAMobileNetX4.__module__.startswith('ofa_b200.') and DynamicSeparableConv2d.__module__.startswith('ofa_b200.')
assert train_one_epoch.__code__.co_filename.startswith('/root/reference/')
assert SRRunManager.__module__ == 'ofa.imagenet_codebase.run_manager.sr_run_manager'
print('ok')
''')
    assert r.returncode == 0 and r.stdout.startswith('ok'), r.stderr[-2000:]


def test_overlay_refuses_to_shadow_an_imported_ofa():
    r = _run('''
import sys, types
sys.modules['ofa'] = types.ModuleType('ofa')
import ofa_b200.compat as c
try:
    c.install_ofa_overlay()
except RuntimeError as e:
    print('ok')
''')
    assert r.returncode == 0 and r.stdout.startswith('ok'), r.stderr[-2000:]
