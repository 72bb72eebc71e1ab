"""OFAMobileNetS4 — the SR *upscaler* supernet (reference ofa/elastic_nn/networks/ofa_mbs4.py:16-178,
base imagenet_codebase/networks/mobilenet_s4.py:15-26).

  x[N,3,H,W] -> 5x5 conv 3->64 + BN                       (dec_first_conv_block)
             -> 4 stages x <=4 elastic MBConv (64ch, ReLU6, identity residual)
             -> 5x5 conv + BN (+ long skip) -> 5x5 conv + BN      (dec_final_conv_blocks)
             -> <=2 x [5x5 conv 64->256 + BN + PixelShuffle(2)]   (blocks[16:], Q1 decides how many)
             -> 5x5 conv 64->3 + BN                               (dec_final_output_conv_block)

Attribute names, block order and state_dict keys equal the reference; the long-skip add is fused
into the epilogue of the conv that precedes it.
"""
import torch
import torch.nn as nn

from ...layers import ConvLayer, IdentityLayer, MobileInvertedResidualBlock
from ...utils import make_divisible, int2list
from ..modules.dynamic_layers import DynamicMBConvLayer
from ..modules.dynamic_op import DynamicBatchNorm2d
from ... import functional as OF
from ... import backend as B
from .supernet_base import ElasticSRSuperNet

__all__ = ['OFAMobileNetS4']


class OFAMobileNetS4(ElasticSRSuperNet):
    _N_STATIC_IN_BLOCKS = 2
    _N_SHUFFLE_GROUPS = 1
    _KS = 5  # kernel size of the static convs

    def __init__(self, bn_param=(0.1, 1e-5), dropout_rate=0.1, base_stage_width=None, width_mult_list=1.0,
                 ks_list=7, expand_ratio_list=6, depth_list=4, pixelshuffle_depth_list=2):
        super().__init__()
        self.width_mult_list = int2list(width_mult_list, 1)
        self.ks_list = int2list(ks_list, 1)
        self.expand_ratio_list = int2list(expand_ratio_list, 1)
        self.depth_list = int2list(depth_list, 1)
        self.pixelshuffle_depth_list = int2list(pixelshuffle_depth_list, 1)
        self.base_stage_width = base_stage_width
        for lst in (self.width_mult_list, self.ks_list, self.expand_ratio_list, self.depth_list,
                    self.pixelshuffle_depth_list):
            lst.sort()

        # stage widths: stem, 4 MBConv stages, 2 tail convs, shuffle conv, output
        stage_width = [64, 64, 64, 64, 64, 64, 64, 256, 3]
        width_list = [[make_divisible(bw * wm, 1) for wm in self.width_mult_list] for bw in stage_width]
        n_mb_stages = 4
        max_depth = max(self.depth_list)
        n_shuffle = max(self.pixelshuffle_depth_list)
        k = self._KS

        first = ConvLayer(3, max(width_list[0]), kernel_size=k, stride=1, act_func=None, use_bn=True)

        self.block_group_info = []
        blocks = []
        feature_dim = width_list[0]
        for stage in range(n_mb_stages):
            output_channel = width_list[1 + stage]
            self.block_group_info.append([len(blocks) + i for i in range(max_depth)])
            for _ in range(max_depth):
                mb = DynamicMBConvLayer(
                    in_channel_list=feature_dim, out_channel_list=output_channel, kernel_size_list=ks_list,
                    expand_ratio_list=expand_ratio_list, stride=1, act_func='relu6', use_se=False,
                )
                blocks.append(MobileInvertedResidualBlock(mb, IdentityLayer(feature_dim, feature_dim)))
                feature_dim = output_channel

        tail = []
        for i in range(2):
            output_channel = width_list[5 + i]
            tail.append(ConvLayer(max(feature_dim), max(output_channel), kernel_size=k, stride=1, act_func=None, use_bn=True))
            feature_dim = output_channel

        self.block_group_info.append([len(blocks) + i for i in range(n_shuffle)])
        for _ in range(n_shuffle):
            blocks.append(ConvLayer(max(feature_dim), max(width_list[7]), kernel_size=k, stride=1,
                                    act_func='pixelshuffle', use_bn=True))
        # attribute order = the reference's registration order (mobilenet_s4.py:21-24), so a seeded
        # init_model() walks the modules identically
        self.blocks = nn.ModuleList(blocks)
        self.dec_first_conv_block = first
        self.dec_final_conv_blocks = nn.ModuleList(tail)

        self.dec_final_output_conv_block = ConvLayer(max(feature_dim), max(width_list[8]), kernel_size=k, stride=1,
                                                     act_func=None, use_bn=True)
        self.dec_final_output_conv_block.out_dtype = torch.float32
        self.dec_final_output_conv_block.out_nchw = True

        self.runtime_depth = [len(block_idx) for block_idx in self.block_group_info]
        self.set_bn_param(momentum=bn_param[0], eps=bn_param[1])

    @staticmethod
    def name():
        return 'OFAMobileNetS4'

    # The planar tensor-core path needs image rows of a multiple of 8 pixels (16-byte TMA row pitch); a frame of any
    # other width would run ~2x slower on the NHWC kernels (16.5 vs 7.9 ms at 956 x 540).  In inference such a frame is
    # computed as TWO column tiles whose widths ARE multiples of 8, overlapping by the receptive field (64 LR pixels):
    # eval-mode BatchNorm is a per-channel affine, so each tile's core columns equal the whole-frame result.
    _SPLIT_HALO = 64
    _SPLIT_MIN_W = 256

    def _column_split(self, x):
        """None, or (core split c, left window end a, right window start b) for a frame that should be computed as two
        column tiles."""
        if not (x.dim() == 4 and x.is_cuda and x.shape[3] % 8 != 0 and x.shape[3] >= self._SPLIT_MIN_W
                and x.shape[2] * x.shape[3] >= 16384 and OF.inference_mode_active(self)
                and OF.get_compute_dtype() != torch.float32 and OF._state['impl'] in (B.IMPL_AUTO, B.IMPL_FAST)
                and not DynamicBatchNorm2d.SET_RUNNING_STATISTICS):
            return None
        w = x.shape[3]
        c = (w // 2) // 8 * 8
        a = (c + self._SPLIT_HALO + 7) // 8 * 8
        b = c - self._SPLIT_HALO
        b -= (8 - (w - b) % 8) % 8               # the right window [b, w) gets a width that is a multiple of 8
        if a > w or b < 0:
            return None
        return c, a, b

    @OF.scoped_forward(OF.pack_plan_signature)
    def forward(self, x):
        split = self._column_split(x)
        if split is not None:
            c, a, b = split
            left = self.forward(x[:, :, :, :a])
            right = self.forward(x[:, :, :, b:])
            s = left.shape[3] // a                # 2 ** (number of PixelShuffle stages that ran)
            out = torch.empty((x.shape[0], left.shape[1], left.shape[2], s * x.shape[3]), dtype=left.dtype,
                              device=left.device)
            out[:, :, :, :s * c] = left[:, :, :, :s * c]
            out[:, :, :, s * c:] = right[:, :, :, s * (c - b):]
            return out
        x = self.dec_first_conv_block(x)
        dec_big_skip = x
        x = self._run_groups(x, 0, 4)
        x = self.dec_final_conv_blocks[0](x, residual=dec_big_skip)   # x = conv(x); x += dec_big_skip
        x = self.dec_final_conv_blocks[1](x)
        x = self._run_groups(x, 4, None)                              # shuffle depth = runtime_depth[0] (Q1)
        return self.dec_final_output_conv_block(x)

    @property
    def module_str(self):
        _str = ''
        for stage_id, block_idx in enumerate(self.block_group_info):
            for idx in block_idx[:self.runtime_depth[stage_id]]:
                _str += self.blocks[idx].module_str + '\n'
        _str += self.dec_first_conv_block.module_str + '\n'
        for block in self.dec_final_conv_blocks:
            _str += block.module_str + '\n'
        return _str + self.dec_final_output_conv_block.module_str + '\n'

    # ---- sub-network selection (ofa_mbs4.py:263-293) ---------------------------------------------
    def _elastic_block_range(self):
        return self.blocks[:-1]   # zipped against len(blocks)-2 settings (Q4)

    def _depth_with_pixel(self, depth, pixel_d):
        depth.insert(-1, pixel_d[0])   # lands BEFORE the last MBConv stage's entry (Q2)
