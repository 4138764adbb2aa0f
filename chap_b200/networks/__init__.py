from .unet import UNet, DualDecoder
from .vnet import VNet, DualDecoder3d
from .net_factory import net_factory
from .net_factory_3d import net_factory_3d
