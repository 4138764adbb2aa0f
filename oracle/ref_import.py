"""TEST INFRASTRUCTURE ONLY -- never imported by the product package `chap_b200`.

Imports the UNMODIFIED reference networks from /root/reference/code so that the
functional restatement in `oracle/nets.py` (and, through the golden fixtures, the
CUDA path) can be pinned against the real thing.  /root/reference exists only in
the build container; on the GPU box `available()` is False and callers fall back
to the committed fixtures under tests/golden/.

The reference modules pull in third-party packages that are not installed here
(fvcore, thop, torchsummary, detectron2, timm) but that the hot-path classes never
use (reference: code/networks/unet.py:12-22, code/networks/vnet.py:4-6,
code/networks/mask2former_transformer_decoder.py:4,15-19).  They are replaced by
empty stub modules in sys.modules; no reference file is touched.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("CHAP_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_CODE, "networks", "unet.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    mod.__dict__.update(attrs)
    if "." in name:
        parent, child = name.rsplit(".", 1)
        if parent not in sys.modules:
            _stub(parent)
        setattr(sys.modules[parent], child, mod)
    return mod


_installed = False


def _install_stubs():
    global _installed
    if _installed:
        return
    import torch

    def _have(modname):
        try:
            __import__(modname)
            return True
        except Exception:
            return False

    if not _have("fvcore.nn.weight_init"):
        _stub("fvcore.nn.weight_init")
    if not _have("thop"):
        _stub("thop", clever_format=None, profile=None)
    if not _have("torchsummary"):
        _stub("torchsummary", summary=None)
    if not _have("detectron2.config"):
        _stub("detectron2.config", configurable=lambda f=None, **k: f)

        class Registry:  # minimal stand-in for detectron2.utils.registry.Registry
            def __init__(self, name):
                self.name = name

            def register(self, obj=None):
                return obj if obj is not None else (lambda o: o)

        _stub("detectron2.utils.registry", Registry=Registry)
    if not _have("timm.models.layers"):
        _stub("timm.models.layers", DropPath=torch.nn.Identity, trunc_normal_tf_=None,
              trunc_normal_=None, to_2tuple=None)
    # evaluators (val_2D / val_3D / test_3D_util) import these at module top
    for name, attrs in (("medpy", {}), ("medpy.metric", {}), ("h5py", {}), ("nibabel", {}),
                        ("SimpleITK", {}), ("skimage", {}), ("skimage.measure", {"label": None})):
        if not _have(name):
            _stub(name, **attrs)
    sys.dont_write_bytecode = True  # reference tree is read-only
    if REF_CODE not in sys.path:
        sys.path.insert(0, REF_CODE)
    _installed = True


def load():
    """Returns a namespace with the reference hot-path classes/functions."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    from networks import unet as r_unet          # code/networks/unet.py
    from networks import vnet as r_vnet          # code/networks/vnet.py
    from networks import FilterDropout as r_fd   # code/networks/FilterDropout.py
    ns = types.SimpleNamespace(
        unet=r_unet, vnet=r_vnet, filter_dropout=r_fd,
        UNet=r_unet.UNet, DualDecoder=r_unet.DualDecoder,
        VNet=r_vnet.VNet, DualDecoder3d=r_vnet.DualDecoder3d,
        perform_dropout=r_fd.perform_dropout,
    )
    return ns


def load_sliding_window():
    """Reference test_3D_util module (code/test_3D_util.py).  Its test_single_case
    hard-codes .cuda() (line 59) so it is only *callable* on a GPU box; here it is
    imported for inspection and for a monkey-patched CPU run in oracle/make_golden.py."""
    _install_stubs()
    import test_3D_util as r_sw
    return r_sw
