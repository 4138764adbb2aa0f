"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product package.

Functional, state-dict driven CPU restatement of the reference networks on the CHAP
hot path, written against plain torch.nn.functional in fp32.  It exists because the
reference tree (/root/reference) cannot travel to the GPU box: there the CUDA path is
compared with THIS file on the same weights/inputs, and this file is pinned against
the real reference classes in the build container (tests/test_oracle_vs_reference.py)
and against the fixtures in tests/golden/ everywhere.

What is restated (reference file:line):
  2D  ConvBlock        code/networks/unet.py:44-60     conv3x3+b -> BN -> LeakyReLU(.01) -> Dropout(p) -> conv3x3+b -> BN -> LeakyReLU
      DownBlock        code/networks/unet.py:63-75     MaxPool2d(2) -> ConvBlock
      UpBlock          code/networks/unet.py:78-99     [conv1x1 -> bilinear x2 align_corners | ConvTranspose k2s2] -> cat(skip, up) -> ConvBlock(p=0)
      Encoder/Decoder  code/networks/unet.py:125-190   5-level pyramid / 4 UpBlocks + out conv3x3
      DualDecoder      code/networks/unet.py:245-292   decoder1 up_type=1 (bilinear), decoder2 ('mcnet') up_type=0 (transposed)
      UNet             code/networks/unet.py:498-552
  3D  ConvBlock        code/networks/vnet.py:8-34      n_stages x (conv3^3+b -> BN3d -> ReLU)
      Downsampling     code/networks/vnet.py:70-94     conv k2 s2 -> BN -> ReLU
      Upsampling       code/networks/vnet.py:97-125    mode 0: ConvTranspose3d k2s2 ; mode 1: trilinear x2 (align_corners) + conv3^3 ; -> BN -> ReLU
      Encoder/Decoder  code/networks/vnet.py:127-223   additive skips, Dropout3d(.5) on x5 / x9 when has_dropout and training
      DualDecoder3d    code/networks/vnet.py:225-238   decoder1 mode 1, decoder2 mode 0
      VNet             code/networks/vnet.py:303-315   encoder + mode-0 decoder

Randomness (nn.Dropout / nn.Dropout3d) is made explicit: `drop` is either None (no
dropout applied: eval mode or p treated as 0), the string "torch" (draw with torch's
generator exactly like the reference modules would), or a dict name->mask holding the
already 1/(1-p)-scaled multiplicative masks (the parity protocol: the same masks are
fed to the CUDA path).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

UNET_CH = (16, 32, 64, 128, 256)
UNET_DROP = (0.05, 0.1, 0.2, 0.3, 0.5)      # code/networks/unet.py:251
BN_EPS = 1e-5
BN_MOM = 0.1


class BNMode:
    """How BatchNorm behaves in one functional pass."""
    def __init__(self, train=True, update_running=True):
        self.train = train
        self.update_running = update_running


def _bn(sd, key, x, mode):
    w, b = sd[key + "weight"], sd[key + "bias"]
    rm, rv = sd[key + "running_mean"], sd[key + "running_var"]
    if not mode.train:
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOM, BN_EPS)
    if mode.update_running:
        out = F.batch_norm(x, rm, rv, w, b, True, BN_MOM, BN_EPS)
        nbt = sd.get(key + "num_batches_tracked")
        if nbt is not None:
            nbt += 1
        return out
    return F.batch_norm(x, None, None, w, b, True, BN_MOM, BN_EPS)


def _drop(x, drop, name, p, channelwise=False):
    if drop is None or p == 0.0:
        return x
    if isinstance(drop, str):
        assert drop == "torch"
        if channelwise:
            return F.dropout3d(x, p, True) if x.dim() == 5 else F.dropout2d(x, p, True)
        return F.dropout(x, p, True)
    m = drop.get(name)
    return x if m is None else x * m


# --------------------------------------------------------------------------- 2D

def _convblock2d(sd, pre, x, mode, drop, p):
    x = F.conv2d(x, sd[pre + "conv_conv.0.weight"], sd[pre + "conv_conv.0.bias"], padding=1)
    x = F.leaky_relu(_bn(sd, pre + "conv_conv.1.", x, mode), 0.01)
    x = _drop(x, drop, pre + "conv_conv.3", p)
    x = F.conv2d(x, sd[pre + "conv_conv.4.weight"], sd[pre + "conv_conv.4.bias"], padding=1)
    return F.leaky_relu(_bn(sd, pre + "conv_conv.5.", x, mode), 0.01)


def unet_encoder(sd, x, mode, drop=None, pre="encoder."):
    feats = [_convblock2d(sd, pre + "in_conv.", x, mode, drop, UNET_DROP[0])]
    for i in range(1, 5):
        h = F.max_pool2d(feats[-1], 2)
        feats.append(_convblock2d(sd, pre + "down%d.maxpool_conv.1." % i, h, mode, drop, UNET_DROP[i]))
    return feats


def unet_decoder(sd, feats, mode, pre, with_features=False):
    bilinear = (pre + "up1.conv1x1.weight") in sd
    x = feats[4]
    for i, skip in zip(range(1, 5), (feats[3], feats[2], feats[1], feats[0])):
        u = pre + "up%d." % i
        if bilinear:
            x = F.conv2d(x, sd[u + "conv1x1.weight"], sd[u + "conv1x1.bias"])
            x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        else:
            x = F.conv_transpose2d(x, sd[u + "up.weight"], sd[u + "up.bias"], stride=2)
        x = _convblock2d(sd, u + "conv.", torch.cat([skip, x], dim=1), mode, None, 0.0)
    out = F.conv2d(x, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"], padding=1)
    return (out, x) if with_features else out


def dualdecoder2d_forward(sd, x, train=True, update_running=True, drop=None, with_feat=False):
    mode = BNMode(train, update_running)
    feats = unet_encoder(sd, x, mode, drop if train else None)
    o1 = unet_decoder(sd, feats, mode, "decoder1.")
    o2 = unet_decoder(sd, feats, mode, "decoder2.")
    return (o1, o2, feats) if with_feat else (o1, o2)


def dualdecoder2d_decode(sd, feats, which, train=True, update_running=True):
    return unet_decoder(sd, feats, BNMode(train, update_running), "decoder%d." % which)


def unet2d_forward(sd, x, train=True, update_running=True, drop=None, with_feats=False):
    mode = BNMode(train, update_running)
    feats = unet_encoder(sd, x, mode, drop if train else None)
    return unet_decoder(sd, feats, mode, "decoder.", with_feats)


# --------------------------------------------------------------------------- 3D

def _convblock3d(sd, pre, x, mode, n_stages):
    for s in range(n_stages):
        x = F.conv3d(x, sd[pre + "conv.%d.weight" % (3 * s)], sd[pre + "conv.%d.bias" % (3 * s)], padding=1)
        x = F.relu(_bn(sd, pre + "conv.%d." % (3 * s + 1), x, mode))
    return x


def _down3d(sd, pre, x, mode):
    x = F.conv3d(x, sd[pre + "conv.0.weight"], sd[pre + "conv.0.bias"], stride=2)
    return F.relu(_bn(sd, pre + "conv.1.", x, mode))


def _up3d(sd, pre, x, mode):
    if (pre + "conv.0.weight") in sd:      # mode_upsampling == 0: ConvTranspose3d is conv.0, BN conv.1
        x = F.conv_transpose3d(x, sd[pre + "conv.0.weight"], sd[pre + "conv.0.bias"], stride=2)
        return F.relu(_bn(sd, pre + "conv.1.", x, mode))
    x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    x = F.conv3d(x, sd[pre + "conv.1.weight"], sd[pre + "conv.1.bias"], padding=1)
    return F.relu(_bn(sd, pre + "conv.2.", x, mode))


def vnet_encoder(sd, x, mode, has_dropout, drop=None, pre="encoder."):
    x1 = _convblock3d(sd, pre + "block_one.", x, mode, 1)
    x2 = _convblock3d(sd, pre + "block_two.", _down3d(sd, pre + "block_one_dw.", x1, mode), mode, 2)
    x3 = _convblock3d(sd, pre + "block_three.", _down3d(sd, pre + "block_two_dw.", x2, mode), mode, 3)
    x4 = _convblock3d(sd, pre + "block_four.", _down3d(sd, pre + "block_three_dw.", x3, mode), mode, 3)
    x5 = _convblock3d(sd, pre + "block_five.", _down3d(sd, pre + "block_four_dw.", x4, mode), mode, 3)
    if has_dropout and mode.train:
        x5 = _drop(x5, drop, pre + "dropout", 0.5, channelwise=True)
    return [x1, x2, x3, x4, x5]


def vnet_decoder(sd, feats, mode, has_dropout, pre, drop=None):
    x1, x2, x3, x4, x5 = feats
    x = _up3d(sd, pre + "block_five_up.", x5, mode) + x4
    x = _convblock3d(sd, pre + "block_six.", x, mode, 3)
    x = _up3d(sd, pre + "block_six_up.", x, mode) + x3
    x = _convblock3d(sd, pre + "block_seven.", x, mode, 3)
    x = _up3d(sd, pre + "block_seven_up.", x, mode) + x2
    x = _convblock3d(sd, pre + "block_eight.", x, mode, 2)
    x = _up3d(sd, pre + "block_eight_up.", x, mode) + x1
    x = _convblock3d(sd, pre + "block_nine.", x, mode, 1)
    if has_dropout and mode.train:
        x = _drop(x, drop, pre + "dropout", 0.5, channelwise=True)
    return F.conv3d(x, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def dualdecoder3d_forward(sd, x, train=True, update_running=True, has_dropout=False, drop=None, with_feat=False):
    mode = BNMode(train, update_running)
    feats = vnet_encoder(sd, x, mode, has_dropout, drop)
    o1 = vnet_decoder(sd, feats, mode, has_dropout, "decoder1.", drop)
    o2 = vnet_decoder(sd, feats, mode, has_dropout, "decoder2.", drop)
    return (o1, o2, feats) if with_feat else (o1, o2)


def dualdecoder3d_decode(sd, feats, which, train=True, update_running=True, has_dropout=False, drop=None):
    return vnet_decoder(sd, feats, BNMode(train, update_running), has_dropout, "decoder%d." % which, drop)


def vnet_forward(sd, x, train=True, update_running=True, has_dropout=False, drop=None):
    mode = BNMode(train, update_running)
    feats = vnet_encoder(sd, x, mode, has_dropout, drop)
    return vnet_decoder(sd, feats, mode, has_dropout, "decoder.", drop)


# ----------------------------------------------------------------- state dicts

def clone_state_dict(sd, requires_grad=False, device="cpu"):
    """Detached fp32 copy (contiguous, standard layout) of a module state_dict, with
    leaf-requires-grad set on floating parameters (not on BN running statistics)."""
    out = OrderedDict()
    for k, v in sd.items():
        t = v.detach().to(device).clone().contiguous()
        if requires_grad and t.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        out[k] = t
    return out
