"""OFAMobileNetX4 — the joint task-aware *downscaler -> upscaler* supernet (reference
ofa/elastic_nn/networks/ofa_mbx4.py:16-254, base imagenet_codebase/networks/mobilenet_x4.py:15-27).

  encoder: <=2 x [3x3 conv + BN + PixelUnshuffle(2)]  (blocks[0:2])
           -> 4 stages x <=4 elastic MBConv -> 3x3 conv+BN (+skip) -> 3x3 conv+BN -> 3x3 conv 64->3 + BN
  decoder: 3x3 conv 3->64 + BN -> 4 stages x <=4 elastic MBConv -> 3x3 conv+BN (+skip) -> 3x3 conv+BN
           -> <=2 x [3x3 conv 64->256 + BN + PixelShuffle(2)] -> 3x3 conv 64->3 + BN
Input and output have the same H x W.
"""
import torch
import torch.nn as nn

from ...layers import ConvLayer, IdentityLayer, MobileInvertedResidualBlock
from ...utils import make_divisible, int2list
from ..modules.dynamic_layers import DynamicMBConvLayer
from ... import functional as OF
from .supernet_base import ElasticSRSuperNet

__all__ = ['OFAMobileNetX4']


class OFAMobileNetX4(ElasticSRSuperNet):
    _N_STATIC_IN_BLOCKS = 4
    _N_SHUFFLE_GROUPS = 2
    _KS = 3

    def __init__(self, bn_param=(0.1, 1e-5), dropout_rate=0.1, base_stage_width=None, width_mult_list=1.0,
                 ks_list=3, expand_ratio_list=6, depth_list=4, pixelshuffle_depth_list=2):
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

        stage_width = [16, 64, 64, 64, 64, 64, 64, 3, 64, 64, 64, 64, 64, 64, 64, 256, 3]
        width_list = [[make_divisible(bw * wm, 1) for wm in self.width_mult_list] for bw in stage_width]
        max_depth = max(self.depth_list)
        n_shuffle = max(self.pixelshuffle_depth_list)
        k = self._KS

        def mb_stages(blocks, feature_dim, widths):
            for output_channel in widths:
                self.block_group_info.append([len(blocks) + i for i in range(max_depth)])
                for _ in range(max_depth):
                    mb = DynamicMBConvLayer(
                        in_channel_list=feature_dim, out_channel_list=output_channel, kernel_size_list=ks_list,
                        expand_ratio_list=expand_ratio_list, stride=1, act_func='relu6', use_se=False,
                    )
                    blocks.append(MobileInvertedResidualBlock(mb, IdentityLayer(feature_dim, feature_dim)))
                    feature_dim = output_channel
            return feature_dim

        def plain_convs(feature_dim, widths):
            layers = []
            for output_channel in widths:
                layers.append(ConvLayer(max(feature_dim), max(output_channel), kernel_size=k, stride=1,
                                        act_func=None, use_bn=True))
                feature_dim = output_channel
            return layers, feature_dim

        # encoder un-shuffle convs (both are always built, whatever max(pixelshuffle_depth_list) is)
        c0 = max(width_list[0])
        blocks = [
            ConvLayer(3, c0, kernel_size=k, stride=1, act_func='pixelunshuffle', use_bn=True),
            ConvLayer(c0 * 4, c0, kernel_size=k, stride=1, act_func='pixelunshuffle', use_bn=True),
        ]
        self.block_group_info = [[0, 1]]
        feature_dim = mb_stages(blocks, width_list[1], width_list[1:5])
        enc_tail, feature_dim = plain_convs(feature_dim, width_list[5:8])

        dec_first = ConvLayer(max(feature_dim), max(width_list[8]), kernel_size=k, stride=1,
                              act_func=None, use_bn=True)
        feature_dim = mb_stages(blocks, width_list[6], width_list[9:13])
        dec_tail, feature_dim = plain_convs(feature_dim, width_list[13:15])

        self.block_group_info.append([len(blocks) + i for i in range(n_shuffle)])
        for _ in range(n_shuffle):
            blocks.append(ConvLayer(max(feature_dim), max(width_list[15]), kernel_size=k, stride=1,
                                    act_func='pixelshuffle', use_bn=True))
        # attribute order = the reference's registration order (mobilenet_x4.py:21-25)
        self.blocks = nn.ModuleList(blocks)
        self.enc_final_conv_blocks = nn.ModuleList(enc_tail)
        self.dec_first_conv_block = dec_first
        self.dec_final_conv_blocks = nn.ModuleList(dec_tail)

        self.dec_final_output_conv_block = ConvLayer(max(feature_dim), max(width_list[16]), kernel_size=k, stride=1,
                                                     act_func=None, use_bn=True)
        self.dec_final_output_conv_block.out_dtype = torch.float32
        self.dec_final_output_conv_block.out_nchw = True

        self.runtime_depth = [len(block_idx) for block_idx in self.block_group_info]
        self.set_bn_param(momentum=bn_param[0], eps=bn_param[1])

    @staticmethod
    def name():
        return 'OFAMobileNetX4'

    @OF.scoped_forward(OF.pack_plan_signature)
    def forward(self, x):
        # encoder
        x = self._run_groups(x, 0, 1)
        enc_big_skip = x
        x = self._run_groups(x, 1, 5)            # depths runtime_depth[0..3] (Q1)
        x = self.enc_final_conv_blocks[0](x, residual=enc_big_skip)
        x = self.enc_final_conv_blocks[1](x)
        x = self.enc_final_conv_blocks[2](x)     # 3-channel learned low-resolution image
        # decoder
        x = self.dec_first_conv_block(x)
        dec_big_skip = x
        x = self._run_groups(x, 5, 9)            # again runtime_depth[0..3] (Q1)
        x = self.dec_final_conv_blocks[0](x, residual=dec_big_skip)
        x = self.dec_final_conv_blocks[1](x)
        x = self._run_groups(x, 9, None)         # shuffle depth = runtime_depth[0]
        return self.dec_final_output_conv_block(x)

    @property
    def module_str(self):
        _str = ''
        for stage_id, block_idx in enumerate(self.block_group_info):
            for idx in block_idx[:self.runtime_depth[stage_id]]:
                _str += self.blocks[idx].module_str + '\n'
        for block in self.enc_final_conv_blocks:
            _str += block.module_str + '\n'
        _str += self.dec_first_conv_block.module_str + '\n'
        for block in self.dec_final_conv_blocks:
            _str += block.module_str + '\n'
        return _str + self.dec_final_output_conv_block.module_str + '\n'

    # ---- sub-network selection (ofa_mbx4.py:345-376) ---------------------------------------------
    def _elastic_block_range(self):
        return self.blocks[2:-2]

    def _depth_with_pixel(self, depth, pixel_d):
        depth.insert(0, pixel_d[0])
        depth.insert(-1, pixel_d[0])
