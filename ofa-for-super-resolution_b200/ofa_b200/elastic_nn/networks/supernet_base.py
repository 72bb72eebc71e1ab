"""Sub-network bookkeeping shared by the two SR supernets.

Everything here is host-side integer logic that must be BIT-EXACT with the reference, including its
observable quirks:
  Q1  forward indexes `runtime_depth` with the position inside the slice of `block_group_info`
  Q2  the pixel-shuffle depth is inserted BEFORE the last entry of the depth list
  Q3  the caller's depth list is mutated in place (so the dict `sample_active_subnet` returns shows it)
  Q4  kernel-size / expand settings are zipped against a block range that may be one longer
  Q6  `random.choice` consumption order: ks per block, e per block, d per stage, pixel_d once
References: ofa/elastic_nn/networks/ofa_mbs4.py:263-370, ofa_mbx4.py:345-453.
"""
import random

import torch

from ...utils import MyNetwork, int2list

_CONSTRAINT_KEYS = {
    'depth': '_depth_include_list',
    'expand_ratio': '_expand_include_list',
    'kernel_size': '_ks_include_list',
    'width_mult': '_width_mult_include_list',
    'pixelshuffle_depth': '_pixelshuffle_depth_include_list',
}


class ElasticSRSuperNet(MyNetwork):
    # filled in by the concrete nets
    _N_STATIC_IN_BLOCKS = 0      # how many entries of self.blocks are NOT elastic MBConv blocks
    _N_SHUFFLE_GROUPS = 0        # groups of block_group_info that hold (un)shuffle ConvLayers

    # ---- forward helper: Q1 ---------------------------------------------------------------------
    def _run_groups(self, x, lo, hi):
        """Run `block_group_info[lo:hi]`.  The depth of the i-th group OF THE SLICE is
        `runtime_depth[i]` — the slice-relative index, exactly as the reference's `enumerate`."""
        for pos, block_idx in enumerate(self.block_group_info[lo:hi]):
            depth = self.runtime_depth[pos]
            for idx in block_idx[:depth]:
                x = self.blocks[idx](x)
        return x

    # ---- elastic MBConv blocks addressed by set_active_subnet -----------------------------------
    def _elastic_block_range(self):
        raise NotImplementedError

    def _depth_with_pixel(self, depth, pixel_d):
        raise NotImplementedError

    def set_active_subnet(self, wid=None, ks=None, e=None, d=None, pixel_d=None):
        n_elastic = len(self.blocks) - self._N_STATIC_IN_BLOCKS
        ks = int2list(ks, n_elastic)
        expand_ratio = int2list(e, n_elastic)
        depth = int2list(d, len(self.block_group_info) - self._N_SHUFFLE_GROUPS)
        pixelshuffle_depth = int2list(pixel_d, self._N_SHUFFLE_GROUPS)
        self._depth_with_pixel(depth, pixelshuffle_depth)  # mutates `depth` in place (Q2, Q3)

        for block, k, ex in zip(self._elastic_block_range(), ks, expand_ratio):
            if k is not None:
                block.mobile_inverted_conv.active_kernel_size = k
            if ex is not None:
                block.mobile_inverted_conv.active_expand_ratio = ex
        for i, dd in enumerate(depth):
            if dd is not None:
                self.runtime_depth[i] = min(len(self.block_group_info[i]), dd)

    def set_output_dtype(self, dtype):
        """Format of the SR image the INFERENCE path returns (the reference API returns fp32; default).  torch.float16
        halves, torch.uint8 quarters the device -> host traffic of a frame: uint8 is exactly what the reference's consumer
        computes from the fp32 tensor, `tensor2img_np` = round(clamp(y, 0, 1) * 255) (sr_run_manager.py:567-597), written
        by the last conv's epilogue.  Training / grad-enabled forwards always return fp32."""
        assert dtype in (torch.float32, torch.float16, torch.uint8)
        self.dec_final_output_conv_block.out_dtype = dtype

    def set_constraint(self, include_list, constraint_type='depth'):
        if constraint_type not in _CONSTRAINT_KEYS:
            raise NotImplementedError
        self.__dict__[_CONSTRAINT_KEYS[constraint_type]] = include_list.copy()

    def clear_constraint(self):
        for key in _CONSTRAINT_KEYS.values():
            self.__dict__[key] = None

    def _candidates(self, key, default):
        got = self.__dict__.get(key, None)
        return default if got is None else got

    def sample_active_subnet(self):
        n_elastic = len(self.blocks) - self._N_STATIC_IN_BLOCKS
        n_depth = len(self.block_group_info) - self._N_SHUFFLE_GROUPS

        def draw(cands, count):
            per_slot = cands if isinstance(cands[0], list) else [cands for _ in range(count)]
            return [random.choice(options) for options in per_slot]

        # order of RNG consumption is part of the contract (Q6)
        ks_setting = draw(self._candidates('_ks_include_list', self.ks_list), n_elastic)
        expand_setting = draw(self._candidates('_expand_include_list', self.expand_ratio_list), n_elastic)
        depth_setting = draw(self._candidates('_depth_include_list', self.depth_list), n_depth)
        pixel_setting = draw(self._candidates('_pixelshuffle_depth_include_list', self.pixelshuffle_depth_list), 1)

        self.set_active_subnet(None, ks_setting, expand_setting, depth_setting, pixel_setting)
        return {'wid': None, 'ks': ks_setting, 'e': expand_setting, 'd': depth_setting, 'pixel_d': pixel_setting}

    # ---- checkpoints -----------------------------------------------------------------------------
    def load_weights_from_net(self, src_model_dict):
        """Load a state_dict saved from either the static (`.conv.weight`, `.bn.`) or the dynamic
        (`.conv.conv.weight`, `.bn.bn.`) layout, with or without DataParallel's `module.` prefix
        (ofa_mbs4.py:221-259)."""
        model_dict = self.state_dict()
        renames = (
            ('.bn.bn.', '.bn.'), ('.conv.conv.weight', '.conv.weight'), ('.linear.linear.', '.linear.'),
            ('.linear.', '.linear.linear.'), ('bn.', 'bn.bn.'), ('conv.weight', 'conv.conv.weight'),
        )
        for src_key, value in src_model_dict.items():
            key = src_key.replace('module.', '') if 'module.' in src_key else src_key
            if key in model_dict:
                new_key = key
            else:
                for old, new in renames:
                    if old in key:
                        new_key = key.replace(old, new)
                        break
                else:
                    raise ValueError(key)
            assert new_key in model_dict, '%s' % new_key
            model_dict[new_key] = value
        self.load_state_dict(model_dict)

    def re_organize_middle_weights(self, expand_ratio_stage=0):
        # the reference walks blocks[2:-2] in BOTH nets (Q8)
        for block in self.blocks[2:-2]:
            block.mobile_inverted_conv.re_organize_middle_weights(expand_ratio_stage)

    # ---- not available in the reference either (Q7): they reference attributes of another net ------
    def get_active_subnet(self, preserve_weight=True):
        raise AttributeError(
            "network-level get_active_subnet is broken in the reference (it reads self.first_conv); use "
            "DynamicMBConvLayer.get_active_subnet per block")

    @staticmethod
    def build_from_config(config):
        raise ValueError('do not support this function')

    def zero_last_gamma(self):
        from ...layers import MobileInvertedResidualBlock, MBInvertedConvLayer, IdentityLayer
        for m in self.modules():
            if isinstance(m, MobileInvertedResidualBlock):
                if isinstance(m.mobile_inverted_conv, MBInvertedConvLayer) and isinstance(m.shortcut, IdentityLayer):
                    m.mobile_inverted_conv.point_linear.bn.weight.data.zero_()
