"""ofa_b200 — B200-native (sm_100a) implementation of the elastic-MBConv super-resolution hot path
of twice154/ofa-for-super-resolution, behind the reference's own `ofa/elastic_nn` module API.

    from ofa_b200.elastic_nn.networks import OFAMobileNetS4, OFAMobileNetX4
    from ofa_b200.elastic_nn.modules import DynamicMBConvLayer, DynamicSeparableConv2d, ...

All activation math runs in libofa_sr_b200.so (include/ofa_sr_b200.h); there is no CPU fallback.
"""
from . import backend, functional  # noqa: F401
from .functional import (set_compute_dtype, get_compute_dtype, set_impl, set_train_dtype, get_train_dtype,  # noqa: F401
                         set_mid_dtype, check_finite, set_overflow_policy, invalidate_packed_weights,
                         set_block_train, set_train_side_stream)

from .graphs import GraphedModule  # noqa: F401,E402

__version__ = '0.2.0'
