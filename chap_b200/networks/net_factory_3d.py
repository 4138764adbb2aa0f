"""String -> 3D model constructor, mirroring code/networks/net_factory_3d.py:7-31 for the branches
on the CHAP hot path ('vnet', 'dualdecoder' x mode 'train'|'test'); unet_3D / attention_unet /
voxresnet are out of scope and return None like an unknown name does."""
from .vnet import VNet, DualDecoder3d


def net_factory_3d(net_type="unet_3D", in_chns=1, class_num=2, mode='train', device="cuda:0", args=None):
    if net_type == "vnet" and mode in ('train', 'test'):
        net = VNet(n_channels=in_chns, n_classes=class_num, normalization='batchnorm',
                   has_dropout=(mode == 'train')).to(device)
    elif net_type == "dualdecoder" and mode in ('train', 'test'):
        net = DualDecoder3d(n_channels=in_chns, n_classes=class_num, normalization='batchnorm',
                            has_dropout=(mode == 'train'), args=args).to(device)
    else:
        net = None
    return net
