"""String -> 2D model constructor, mirroring code/networks/net_factory.py:11-24 for the branches that
are on the CHAP hot path ('unet', 'dualdecoder'); the other reference branches (unetp, dual_student,
resunet, EfficientNet/Swin baselines) are out of scope and return None like an unknown name does."""
from .unet import UNet, DualDecoder


def net_factory(net_type="unet", in_chns=1, class_num=3, device="cuda:0", args=None):
    if net_type == "unet":
        net = UNet(in_chns=in_chns, class_num=class_num).to(device)
    elif net_type == "dualdecoder":
        net = DualDecoder(in_chns=in_chns, class_num=class_num, args=args).to(device)
    else:
        net = None
    return net
