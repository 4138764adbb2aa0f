"""3D V-Net family of the CHAP hot path on the sm_100a kernels.

Drop-in for the reference classes of code/networks/vnet.py (ConvBlock :8-34, DownsamplingConvBlock
:70-94, Upsampling_function :97-125, Encoder :127-168, Decoder :170-223, DualDecoder3d :225-238,
VNet :303-315): same constructors, forward signatures, attribute names and state-dict keys.  The
factories only ever build normalization='batchnorm' without residual blocks
(code/networks/net_factory_3d.py:16-27); other settings are outside the hot path and raise.
The nn.* children are parameter holders (created in the reference's order); forwards run the fused
kernels of libchap_b200 on channels-last (NDHWC) fp32 tensors.
"""
import torch
import torch.nn as nn

from .. import ops
from .._lib import CONV_DOWN2, CONV_K1, CONV_K3, CONV_UP2


def _check_norm(normalization):
    if normalization != 'batchnorm':
        raise NotImplementedError("only normalization='batchnorm' is on the CHAP hot path (got %r)" % (normalization,))


def _conv_bn_relu(x, conv, bn, kind, residual=None, drop_nc=None):
    if drop_nc is None and ops.eval_fusable(bn):       # inference: BatchNorm, ReLU and the skip add run in the conv epilogue
        return ops.conv_bn_act_eval(x, conv.weight, conv.bias, kind, bn, 0.0, residual=residual)
    y, sums = ops.conv_stats(x, conv.weight, conv.bias, kind, bn.training, feeds_train_bn=bn.training, bn=bn)
    return ops.bn_act(y, bn, 0.0, sums=sums, residual=residual, drop_nc=drop_nc)


class ConvBlock(nn.Module):
    """n_stages x (conv3^3 -> BN -> ReLU)  (vnet.py:8-34)"""

    def __init__(self, n_stages, n_filters_in, n_filters_out, normalization='none'):
        super().__init__()
        _check_norm(normalization)
        ops_ = []
        for i in range(n_stages):
            ops_.append(nn.Conv3d(n_filters_in if i == 0 else n_filters_out, n_filters_out, 3, padding=1))
            ops_.append(nn.BatchNorm3d(n_filters_out))
            ops_.append(nn.ReLU(inplace=True))
        self.conv = nn.Sequential(*ops_)
        self.n_stages = n_stages

    def forward(self, x, drop_nc=None):
        """drop_nc: optional [N, C] Dropout3d factor fused into the LAST stage's epilogue."""
        for s in range(self.n_stages):
            last = s == self.n_stages - 1
            x = _conv_bn_relu(x, self.conv[3 * s], self.conv[3 * s + 1], CONV_K3, drop_nc=drop_nc if last else None)
        return x


class DownsamplingConvBlock(nn.Module):
    """conv k2 s2 -> BN -> ReLU  (vnet.py:70-94)"""

    def __init__(self, n_filters_in, n_filters_out, stride=2, normalization='none'):
        super().__init__()
        _check_norm(normalization)
        assert stride == 2
        self.conv = nn.Sequential(nn.Conv3d(n_filters_in, n_filters_out, stride, padding=0, stride=stride),
                                  nn.BatchNorm3d(n_filters_out), nn.ReLU(inplace=True))

    def forward(self, x):
        return _conv_bn_relu(x, self.conv[0], self.conv[1], CONV_DOWN2)


class Upsampling_function(nn.Module):
    """mode 0: ConvTranspose3d k2 s2; mode 1: trilinear x2 (align_corners) + conv3^3; then BN -> ReLU
    (vnet.py:97-125).  `skip` (the encoder feature added right after, vnet.py:202-215) is fused into
    the BN/ReLU epilogue."""

    def __init__(self, n_filters_in, n_filters_out, stride=2, normalization='none', mode_upsampling=1):
        super().__init__()
        _check_norm(normalization)
        assert stride == 2
        self.mode_upsampling = mode_upsampling
        layers = []
        if mode_upsampling == 0:
            layers.append(nn.ConvTranspose3d(n_filters_in, n_filters_out, stride, padding=0, stride=stride))
        elif mode_upsampling == 1:
            layers.append(nn.Upsample(scale_factor=stride, mode="trilinear", align_corners=True))
            layers.append(nn.Conv3d(n_filters_in, n_filters_out, kernel_size=3, padding=1))
        else:
            raise NotImplementedError("mode_upsampling 2 (nearest) is outside the CHAP hot path")
        layers.append(nn.BatchNorm3d(n_filters_out))
        layers.append(nn.ReLU(inplace=True))
        self.conv = nn.Sequential(*layers)

    def forward(self, x, skip=None):
        if self.mode_upsampling == 0:
            return _conv_bn_relu(x, self.conv[0], self.conv[1], CONV_UP2, residual=skip)
        return _conv_bn_relu(ops.upsample2x(x), self.conv[1], self.conv[2], CONV_K3, residual=skip)


def _dropout3d_factor(x, training):
    """nn.Dropout3d(0.5) as an [N, C] factor (None when inactive)."""
    if not training:
        return None
    return torch.empty(x.shape[0], x.shape[1], device=x.device).bernoulli_(0.5).mul_(2.0)


class Encoder(nn.Module):
    """vnet.py:127-168"""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization='none', has_dropout=False,
                 has_residual=False):
        super().__init__()
        if has_residual:
            raise NotImplementedError("has_residual is never set by the reference factories")
        self.has_dropout = has_dropout
        nf, norm = n_filters, normalization
        self.block_one = ConvBlock(1, n_channels, nf, normalization=norm)
        self.block_one_dw = DownsamplingConvBlock(nf, 2 * nf, normalization=norm)
        self.block_two = ConvBlock(2, nf * 2, nf * 2, normalization=norm)
        self.block_two_dw = DownsamplingConvBlock(nf * 2, nf * 4, normalization=norm)
        self.block_three = ConvBlock(3, nf * 4, nf * 4, normalization=norm)
        self.block_three_dw = DownsamplingConvBlock(nf * 4, nf * 8, normalization=norm)
        self.block_four = ConvBlock(3, nf * 8, nf * 8, normalization=norm)
        self.block_four_dw = DownsamplingConvBlock(nf * 8, nf * 16, normalization=norm)
        self.block_five = ConvBlock(3, nf * 16, nf * 16, normalization=norm)
        self.dropout = nn.Dropout3d(p=0.5, inplace=False)

    def forward(self, input, drop_nc=None):
        """drop_nc: optional explicit [N, 16*nf] Dropout3d factor for x5 (parity protocol)."""
        x1 = self.block_one(input)
        x2 = self.block_two(self.block_one_dw(x1))
        x3 = self.block_three(self.block_two_dw(x2))
        x4 = self.block_four(self.block_three_dw(x3))
        x4_dw = self.block_four_dw(x4)
        if self.has_dropout and drop_nc is None:
            drop_nc = _dropout3d_factor(x4_dw, self.training)
        x5 = self.block_five(x4_dw, drop_nc if self.has_dropout else None)
        return [x1, x2, x3, x4, x5]


class Decoder(nn.Module):
    """vnet.py:170-223"""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization='none', has_dropout=False,
                 has_residual=False, up_type=0):
        super().__init__()
        if has_residual:
            raise NotImplementedError("has_residual is never set by the reference factories")
        self.has_dropout = has_dropout
        nf, norm = n_filters, normalization
        self.block_five_up = Upsampling_function(nf * 16, nf * 8, normalization=norm, mode_upsampling=up_type)
        self.block_six = ConvBlock(3, nf * 8, nf * 8, normalization=norm)
        self.block_six_up = Upsampling_function(nf * 8, nf * 4, normalization=norm, mode_upsampling=up_type)
        self.block_seven = ConvBlock(3, nf * 4, nf * 4, normalization=norm)
        self.block_seven_up = Upsampling_function(nf * 4, nf * 2, normalization=norm, mode_upsampling=up_type)
        self.block_eight = ConvBlock(2, nf * 2, nf * 2, normalization=norm)
        self.block_eight_up = Upsampling_function(nf * 2, nf, normalization=norm, mode_upsampling=up_type)
        self.block_nine = ConvBlock(1, nf, nf, normalization=norm)
        self.out_conv = nn.Conv3d(nf, n_classes, 1, padding=0)
        self.dropout = nn.Dropout3d(p=0.5, inplace=False)

    def forward(self, features, drop_nc=None):
        x1, x2, x3, x4, x5 = features[0], features[1], features[2], features[3], features[4]
        x = self.block_six(self.block_five_up(x5, skip=x4))
        x = self.block_seven(self.block_six_up(x, skip=x3))
        x = self.block_eight(self.block_seven_up(x, skip=x2))
        x = self.block_eight_up(x, skip=x1)
        if self.has_dropout and drop_nc is None:
            drop_nc = _dropout3d_factor(x, self.training)
        x9 = self.block_nine(x, drop_nc if self.has_dropout else None)
        return ops.conv(x9, self.out_conv.weight, self.out_conv.bias, CONV_K1)


class DualDecoder3d(nn.Module):
    """Shared encoder, decoder1 = trilinear upsampling, decoder2 = transposed conv  (vnet.py:225-238)"""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization='none', has_dropout=False,
                 has_residual=False, args=None):
        super().__init__()
        self.encoder = Encoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual)
        self.decoder1 = Decoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual, 1)
        self.decoder2 = Decoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual, 0)

    def forward(self, input):
        features = self.encoder(input)
        return self.decoder1(features), self.decoder2(features)


class VNet(nn.Module):
    """Encoder + transposed-conv decoder  (vnet.py:303-315)"""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization='none', has_dropout=False,
                 has_residual=False):
        super().__init__()
        self.encoder = Encoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual)
        self.decoder = Decoder(n_channels, n_classes, n_filters, normalization, has_dropout, has_residual, 0)

    def forward(self, input):
        return self.decoder(self.encoder(input))
