"""`ofa` import shim: makes the reference's import paths resolve to the B200 modules.

    import ofa_b200.compat as compat
    compat.install_ofa_overlay()                          # no reference checkout: `ofa.*` = the hot-path modules only
    compat.install_ofa_overlay('/path/to/reference')      # with one: everything the hot path does NOT own (run managers,
                                                          # data providers, the training loop, model zoo) is the reference's

After the call `from ofa.elastic_nn.networks import OFAMobileNetS4`, `from ofa.elastic_nn.modules.dynamic_op import
DynamicSeparableConv2d`, `from ofa.layers import ConvLayer` ... (train_ofa_net_sr_simple.py:13-14, eval_ofa_net_sr.py:13-15,
progressive_shrinking.py:16-21) import the drop-in classes, so the reference's scripts and loops run on the CUDA path
unchanged.  The overlay works through `sys.modules`: the hot-path module names are bound to `ofa_b200`'s modules BEFORE
anything imports them, and the synthetic `ofa` / `ofa.elastic_nn` packages keep the reference directories on their
`__path__` for every other sub-module.
"""
import importlib
import os
import sys
import types

# reference module name -> ofa_b200 module name (the boundary is the module API)
_HOT_PATH = {
    'ofa.elastic_nn.modules': 'ofa_b200.elastic_nn.modules',
    'ofa.elastic_nn.modules.dynamic_op': 'ofa_b200.elastic_nn.modules.dynamic_op',
    'ofa.elastic_nn.modules.dynamic_layers': 'ofa_b200.elastic_nn.modules.dynamic_layers',
    'ofa.elastic_nn.networks': 'ofa_b200.elastic_nn.networks',
    'ofa.elastic_nn.networks.ofa_mbs4': 'ofa_b200.elastic_nn.networks.ofa_mbs4',
    'ofa.elastic_nn.networks.ofa_mbx4': 'ofa_b200.elastic_nn.networks.ofa_mbx4',
    'ofa.elastic_nn.utils': 'ofa_b200.elastic_nn.utils',
    'ofa.layers': 'ofa_b200.layers',
}
# without a reference checkout these resolve to the B200 restatements as well
_STANDALONE = {
    'ofa.utils': 'ofa_b200.utils',
    'ofa.elastic_nn.training': 'ofa_b200.elastic_nn.training',
    'ofa.elastic_nn.training.progressive_shrinking': 'ofa_b200.elastic_nn.training.progressive_shrinking',
}


def _package(name, path):
    m = types.ModuleType(name)
    m.__path__ = list(path)
    m.__package__ = name
    sys.modules[name] = m
    return m


def install_ofa_overlay(reference_root=None):
    """Bind the `ofa.*` names.  `reference_root` = directory that CONTAINS the reference's `ofa/` package (optional).
    Must run before anything imported `ofa`; returns the synthetic `ofa` package."""
    already = [k for k in sys.modules if k == 'ofa' or k.startswith('ofa.')]
    if already and not getattr(sys.modules.get('ofa'), '_ofa_b200_overlay', False):
        raise RuntimeError('`ofa` was imported before install_ofa_overlay(): %s' % already[:3])
    ref_pkg = os.path.join(reference_root, 'ofa') if reference_root else None
    if ref_pkg and not os.path.isdir(ref_pkg):
        raise FileNotFoundError(ref_pkg)
    ofa = _package('ofa', [ref_pkg] if ref_pkg else [])
    ofa._ofa_b200_overlay = True
    enn = _package('ofa.elastic_nn', [os.path.join(ref_pkg, 'elastic_nn')] if ref_pkg else [])
    ofa.elastic_nn = enn
    table = dict(_HOT_PATH)
    if not ref_pkg:
        table.update(_STANDALONE)
    for ref_name, our_name in sorted(table.items()):
        mod = importlib.import_module(our_name)
        sys.modules[ref_name] = mod
        parent, _, leaf = ref_name.rpartition('.')
        setattr(sys.modules[parent], leaf, mod)
    return ofa
