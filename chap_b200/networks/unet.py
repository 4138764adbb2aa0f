"""2D U-Net family of the CHAP hot path on the sm_100a kernels.

Drop-in for the reference classes of code/networks/unet.py (ConvBlock :44-60, DownBlock :63-75,
UpBlock :78-99, Encoder :125-151, Decoder :153-190, DualDecoder :245-292, UNet :498-552):
same constructor signatures, forward signatures, attribute names and state-dict keys/shapes.
The nn.Conv2d / nn.BatchNorm2d / ... children are *parameter holders* created in the reference's
order (so a given torch seed yields the reference's initial weights and reference checkpoints load
with load_state_dict); their own forward is never called -- every forward below runs the fused
kernels of libchap_b200 through chap_b200.ops on channels-last fp32 tensors.
"""
import torch
import torch.nn as nn

from .. import ops
from .._lib import CONV_K1, CONV_K3, CONV_UP2
from .FilterDropout import perform_dropout

LEAKY_SLOPE = 0.01          # nn.LeakyReLU() default used by the reference (unet.py:52,56)


def _elementwise_dropout_mask(like, p, training):
    """nn.Dropout(p) as a multiplicative 1/(1-p)-scaled mask (None when inactive)."""
    if not training or p <= 0.0:
        return None
    keep = 1.0 - p
    return torch.empty_like(like).bernoulli_(keep).div_(keep)


class ConvBlock(nn.Module):
    """conv3x3 -> BN -> LeakyReLU -> Dropout(p) -> conv3x3 -> BN -> LeakyReLU  (unet.py:44-60)"""

    def __init__(self, in_channels, out_channels, dropout_p):
        super().__init__()
        layers = [nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
                  nn.LeakyReLU(), nn.Dropout(dropout_p),
                  nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
                  nn.LeakyReLU()]
        self.conv_conv = nn.Sequential(*layers)
        self.dropout_p = dropout_p

    def forward(self, x, drop_mask=None, cat=None):
        """drop_mask: optional explicit (already 1/(1-p) scaled) dropout mask -- the parity-test
        protocol; by default the mask is drawn like nn.Dropout would.
        cat: the block runs on torch.cat([x, cat], dim=1); the concat is fused into the first convolution."""
        c1, b1, _, _, c2, b2, _ = self.conv_conv
        if drop_mask is None and ops.eval_fusable(b1) and ops.eval_fusable(b2):      # inference (dropout inactive): conv + BatchNorm + LeakyReLU per kernel
            a = ops.conv_bn_act_eval(x, c1.weight, c1.bias, CONV_K3, b1, LEAKY_SLOPE, cat=cat)
            return ops.conv_bn_act_eval(a, c2.weight, c2.bias, CONV_K3, b2, LEAKY_SLOPE)
        y, sums = ops.conv_stats(x, c1.weight, c1.bias, CONV_K3, b1.training, feeds_train_bn=b1.training, cat=cat, bn=b1)
        rng = None
        if drop_mask is None and self.training:
            rng = ops.dropout_rng(self.dropout_p, y.shape[1])          # nn.Dropout(p) with the mask generated inside the kernels
            if rng is None:
                drop_mask = _elementwise_dropout_mask(y, self.dropout_p, self.training)
        a = ops.bn_act(y, b1, LEAKY_SLOPE, sums=sums, drop_el=drop_mask, drop_rng=rng)
        y, sums = ops.conv_stats(a, c2.weight, c2.bias, CONV_K3, b2.training, feeds_train_bn=b2.training, bn=b2)
        return ops.bn_act(y, b2, LEAKY_SLOPE, sums=sums)


class DownBlock(nn.Module):
    """MaxPool2d(2) -> ConvBlock  (unet.py:63-75)"""

    def __init__(self, in_channels, out_channels, dropout_p):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), ConvBlock(in_channels, out_channels, dropout_p))

    def forward(self, x, drop_mask=None):
        return self.maxpool_conv[1](ops.maxpool2(x), drop_mask)


class UpBlock(nn.Module):
    """[conv1x1 -> bilinear x2 | ConvTranspose2d k2 s2] -> cat(skip, up) -> ConvBlock  (unet.py:78-99)"""

    def __init__(self, in_channels1, in_channels2, out_channels, dropout_p, bilinear=True):
        super().__init__()
        self.bilinear = bilinear
        if bilinear:
            self.conv1x1 = nn.Conv2d(in_channels1, in_channels2, kernel_size=1)
            self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        else:
            self.up = nn.ConvTranspose2d(in_channels1, in_channels2, kernel_size=2, stride=2)
        self.conv = ConvBlock(in_channels2 * 2, out_channels, dropout_p)

    def forward(self, x1, x2):
        if self.bilinear:
            x1 = ops.upsample2x(ops.conv(x1, self.conv1x1.weight, self.conv1x1.bias, CONV_K1))
        else:
            x1 = ops.conv(x1, self.up.weight, self.up.bias, CONV_UP2)
        return self.conv(x2, cat=x1)


class Encoder(nn.Module):
    """5-level feature pyramid  (unet.py:125-151)"""

    def __init__(self, params):
        super().__init__()
        self.params = params
        self.in_chns = params['in_chns']
        self.ft_chns = params['feature_chns']
        self.n_class = params['class_num']
        self.dropout = params['dropout']
        assert len(self.ft_chns) == 5
        ch, dp = self.ft_chns, self.dropout
        self.in_conv = ConvBlock(self.in_chns, ch[0], dp[0])
        self.down1 = DownBlock(ch[0], ch[1], dp[1])
        self.down2 = DownBlock(ch[1], ch[2], dp[2])
        self.down3 = DownBlock(ch[2], ch[3], dp[3])
        self.down4 = DownBlock(ch[3], ch[4], dp[4])

    def forward(self, x, drop_masks=None):
        m = drop_masks if drop_masks is not None else [None] * 5
        x0 = self.in_conv(x, m[0])
        x1 = self.down1(x0, m[1])
        x2 = self.down2(x1, m[2])
        x3 = self.down3(x2, m[3])
        x4 = self.down4(x3, m[4])
        return [x0, x1, x2, x3, x4]


class Decoder(nn.Module):
    """4 UpBlocks + conv3x3 to class logits  (unet.py:153-190)"""

    def __init__(self, params):
        super().__init__()
        self.params = params
        self.in_chns = params['in_chns']
        self.ft_chns = params['feature_chns']
        self.n_class = params['class_num']
        self.bilinear = params['up_type']
        assert len(self.ft_chns) == 5
        ch, bil = self.ft_chns, bool(self.bilinear)
        self.up1 = UpBlock(ch[4], ch[3], ch[3], dropout_p=0.0, bilinear=bil)
        self.up2 = UpBlock(ch[3], ch[2], ch[2], dropout_p=0.0, bilinear=bil)
        self.up3 = UpBlock(ch[2], ch[1], ch[1], dropout_p=0.0, bilinear=bil)
        self.up4 = UpBlock(ch[1], ch[0], ch[0], dropout_p=0.0, bilinear=bil)
        self.out_conv = nn.Conv2d(ch[0], self.n_class, kernel_size=3, padding=1)

    def forward(self, feature, with_features=False):
        x0, x1, x2, x3, x4 = feature[0], feature[1], feature[2], feature[3], feature[4]
        x = self.up1(x4, x3)
        x = self.up2(x, x2)
        x = self.up3(x, x1)
        x = self.up4(x, x0)
        output = ops.conv(x, self.out_conv.weight, self.out_conv.bias, CONV_K3)
        return (output, x) if with_features else output


def _unet_params(in_chns, class_num, up_type):
    return {'in_chns': in_chns, 'feature_chns': [16, 32, 64, 128, 256],
            'dropout': [0.05, 0.1, 0.2, 0.3, 0.5], 'class_num': class_num,
            'up_type': up_type, 'acti_func': 'relu'}


class DualDecoder(nn.Module):
    """Shared encoder + two decoders  (unet.py:245-292).  args["decoder_type"]: 'same' (second bilinear
    decoder) or 'mcnet' (transposed-conv decoder, the CHAP default)."""

    def __init__(self, in_chns, class_num, args):
        super().__init__()
        self.encoder = Encoder(_unet_params(in_chns, class_num, 1))
        self.decoder1 = Decoder(_unet_params(in_chns, class_num, 1))
        self.decoder_type = args["decoder_type"]
        if self.decoder_type == 'same':
            self.decoder2 = Decoder(_unet_params(in_chns, class_num, 1))
        elif self.decoder_type == 'mcnet':
            self.decoder2 = Decoder(_unet_params(in_chns, class_num, 0))
        else:
            raise NotImplementedError("decoder_type %r is outside the CHAP hot path (use 'mcnet' or 'same')"
                                      % (self.decoder_type,))

    def forward(self, x, with_feat=False, dropout=False, dropout_level=None, scores=None, comp_dropout=False, dropout_masks=None):
        """Reference signature (unet.py:277); `dropout_masks` is a trailing extension: explicit per-level channel factors for
        perform_dropout (the parity-test protocol), ignored unless dropout=True."""
        feature = self.encoder(x)
        if dropout:
            feature1, feature2 = perform_dropout(feature, dropout_level, scores, comp_dropout, masks=dropout_masks)
            output1 = self.decoder1(feature1)
            output2 = self.decoder2(feature2)
        else:
            output1 = self.decoder1(feature)
            output2 = self.decoder2(feature)
        if with_feat:
            return output1, output2, feature
        return output1, output2


class UNet(nn.Module):
    """Encoder + one bilinear decoder  (unet.py:498-552)."""

    def __init__(self, in_chns, class_num):
        super().__init__()
        self.encoder = Encoder(_unet_params(in_chns, class_num, 1))
        self.decoder = Decoder(_unet_params(in_chns, class_num, 1))

    def forward(self, x, with_feats=False):
        return self.decoder(self.encoder(x), with_feats)

    def foward_encoder(self, x):          # [sic] -- the reference's spelling (unet.py:521)
        return self.encoder(x)

    def forward_decoder(self, x, dropout_flag=False, dropout_level=None):
        if self.training:
            x = self.perform_dropout(x, dropout_flag, dropout_level)
        return self.decoder(x)

    def perform_dropout(self, x, dropout_flag=False, level=None):
        """unet.py:532-552: append a Dropout2d(0.5)-ed copy of the unlabelled half to the batch."""
        out = []
        for idx, feat in enumerate(x):
            half = feat.shape[0] // 2
            unlab = feat[half:]
            if dropout_flag and idx in level:
                keep = torch.empty(unlab.shape[0], unlab.shape[1], device=feat.device).bernoulli_(0.5).mul_(2.0)
                unlab = ops.channel_scale(unlab, keep)
            out.append(torch.cat((feat, unlab)))
        return out
