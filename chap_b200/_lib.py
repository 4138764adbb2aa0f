"""ctypes binding of libchap_b200.so (the C ABI declared in include/chap_b200.h).

The product path has no CPU fallback: if the shared library is missing or an entry point
fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHAP_B200_LIB", os.path.join(_HERE, "lib", "libchap_b200.so"))   # override: developer builds (debug hooks)


class BnTrainArgs(ctypes.Structure):          # chap_bn_train_args
    _fields_ = [("gamma", c_void_p), ("beta", c_void_p), ("eps", c_float), ("momentum", c_float),
                ("running_mean", c_void_p), ("running_var", c_void_p), ("num_batches_tracked", c_void_p),
                ("mean_invstd", c_void_p), ("scale_shift", c_void_p), ("stats_persistent", c_int32), ("reserved_", c_int32)]


class DropoutRng(ctypes.Structure):           # chap_dropout_rng
    _fields_ = [("p", c_float), ("seed", ctypes.c_uint64), ("subsequence", ctypes.c_uint64), ("epoch_dev", c_void_p)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("nd", c_int32), ("n", c_int32), ("in_d", c_int32), ("in_h", c_int32),
                ("in_w", c_int32), ("cin", c_int32), ("cout", c_int32)]


class PackItem(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("w_fwd", c_void_p), ("w_dgrad", c_void_p), ("desc", ConvDesc)]


class Level(ctypes.Structure):
    _fields_ = [("g", c_void_p), ("f", c_void_p), ("out", c_void_p), ("rows", c_int64), ("c", c_int32),
                ("pad_", c_int32)]


class SwDesc(ctypes.Structure):
    _fields_ = [("vol", c_int32 * 3), ("patch", c_int32 * 3), ("nwin", c_int32 * 3), ("stride", c_int32 * 3),
                ("c", c_int32), ("pad_", c_int32)]


CONV_K3, CONV_K1, CONV_DOWN2, CONV_UP2 = 0, 1, 2, 3
LABEL_I64, LABEL_F32 = 0, 1
DIST_KL, DIST_DICE = 0, 1
STAT_SLOTS = 16
PERTURB_MODES = {"sample": 0, "channel": 1, "spatial": 2, "channel_spatial": 3}

P = c_void_p
I, L, F = c_int32, c_int64, c_float
_CD, _SW = POINTER(ConvDesc), POINTER(SwDesc)

# name -> (restype, argtypes); every symbol declared in include/chap_b200.h
SIGNATURES = {
    "chap_last_error": (c_char_p, []),
    "chap_abi_version": (I, []),
    "chap_check_device": (I, []),
    "chap_launch_count": (c_uint64, []),
    "chap_reset_launch_count": (None, []),
    "chap_timing_enable": (None, [I]),
    "chap_timing_report": (I, [ctypes.c_char_p, c_size_t]),
    "chap_set_force_simt": (None, [I]),
    "chap_get_force_simt": (I, []),
    "chap_set_pdl": (None, [I]),
    "chap_get_pdl": (I, []),
    "chap_set_conv_precision": (None, [I]),
    "chap_get_conv_precision": (I, []),
    "chap_conv_packed_elems": (c_size_t, [_CD]),
    "chap_conv_pack_weights": (I, [_CD, P, P, P, P]),
    "chap_conv_pack_weights_batched": (I, [POINTER(PackItem), I, P]),
    "chap_conv_fwd": (I, [_CD, P, P, P, P, P, P]),
    "chap_conv_bn_fwd": (I, [_CD, P, P, P, P, P, P, P]),
    "chap_conv_bn_act_fwd": (I, [_CD, P, P, P, P, F, P, P, P]),
    "chap_conv_dgrad": (I, [_CD, P, P, P, P]),
    "chap_conv_dgrad_split_supported": (I, [_CD, I]),
    "chap_conv_dgrad_split": (I, [_CD, P, P, P, I, P, P]),
    "chap_conv_wgrad_workspace_bytes": (c_size_t, [_CD]),
    "chap_conv_wgrad": (I, [_CD, P, P, P, P, P, c_size_t, P]),
    "chap_conv_wgrad_acc": (I, [_CD, P, P, P, P, P, c_size_t, P, P]),
    "chap_bn_act_bwd_acc": (I, [P, P, P, P, F, P, P, I, L, I, I, P, I, P, P, P, P]),
    "chap_channel_stats": (I, [P, L, I, P, P]),
    "chap_bn_finalize": (I, [P, I, L, P, P, F, F, P, P, P, P, P, I, P]),
    "chap_bn_eval_params": (I, [P, P, P, P, F, P, P, I, P]),
    "chap_bn_act_fwd": (I, [P, P, F, P, P, P, I, L, I, P, P]),
    "chap_bn_act_bwd": (I, [P, P, P, P, P, F, P, P, I, L, I, I, P, P, P, P, P]),
    "chap_bn_act_fwd_rng": (I, [P, P, F, P, ctypes.POINTER(DropoutRng), P, I, L, I, P, P]),
    "chap_bn_act_bwd_rng": (I, [P, P, P, P, F, P, ctypes.POINTER(DropoutRng), I, L, I, I, P, I, P, P, P, I, P]),
    "chap_maxpool2_fwd": (I, [P, I, I, I, I, P, P]),
    "chap_maxpool2_bwd": (I, [P, P, I, I, I, I, P, P]),
    "chap_upsample2x_fwd": (I, [P, I, I, I, I, I, I, P, P]),
    "chap_upsample2x_bwd": (I, [P, I, I, I, I, I, I, P, P]),
    "chap_concat_channels": (I, [P, P, L, I, I, P, P]),
    "chap_split_channels": (I, [P, L, I, I, P, P, P]),
    "chap_channel_scale": (I, [P, P, I, L, I, P, P]),
    "chap_feature_dropout_fwd": (I, [P, P, P, I, I, L, I, P, P, P]),
    "chap_feature_dropout_bwd": (I, [P, P, P, P, I, I, L, I, P, P]),
    "chap_axpy": (I, [P, P, F, L, P, P]),
    "chap_mask_mix": (I, [P, P, P, I, L, I, P, P]),
    "chap_pseudo_label": (I, [P, P, L, I, P, P, P, P, P, P]),
    "chap_softmax": (I, [P, L, I, P, P]),
    "chap_argmax": (I, [P, P, L, I, P, P]),
    "chap_dice_ce_fwd": (I, [P, P, I, P, I, I, L, I, P, P]),
    "chap_dice_ce_bwd": (I, [P, P, I, P, I, I, L, I, P, I, P, P]),
    "chap_mix_loss_finalize": (I, [P, P, I, F, F, P, P]),
    "chap_mix_loss_coef": (I, [P, P, I, F, F, P, P, P, P]),
    "chap_consistency_fwd": (I, [P, P, P, I, L, I, P, P]),
    "chap_consistency_bwd": (I, [P, P, P, I, L, I, P, P, P]),
    "chap_patch_score": (I, [P, P, P, I, I, I, I, I, I, P, P]),
    "chap_patch_mask": (I, [P, P, I, I, I, I, I, I, P, P]),
    "chap_largest_cc_workspace_bytes": (c_size_t, [I, I, I, I, I]),
    "chap_largest_cc": (I, [P, I, I, I, I, I, I, P, P, c_size_t, P]),
    "chap_perturb_workspace_elems": (c_size_t, [POINTER(Level), I, I]),
    "chap_perturb_fwd": (I, [POINTER(Level), I, I, I, F, F, P, c_size_t, P]),
    "chap_l2n_sample_axpy": (I, [P, P, F, I, L, P, P, P]),
    "chap_l2n_sample_axpy_batched": (I, [POINTER(Level), I, I, F, P, P]),
    "chap_sgd_momentum_lrdev": (I, [P, P, P, L, P, F, F, F, P]),
    "chap_sgd_momentum": (I, [P, P, P, L, F, F, F, F, I, P]),
    "chap_schedule_step": (I, [P, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, L, P, P, P]),
    "chap_gather2d": (I, [P, I, P, P, I, I, I, I, I, P, P]),
    "chap_label_overlap": (I, [P, P, L, I, P, P]),
    "chap_sw_extract": (I, [_SW, P, I, I, P, P]),
    "chap_sw_aggregate": (I, [_SW, P, I, P, P, P, P]),
}

_lib = None


def load():
    """Load the shared library (no GPU needed for loading) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libchap_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C chap_b200/csrc`; chap_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.chap_abi_version() != 1:
        raise RuntimeError("libchap_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().chap_last_error()
        raise RuntimeError("libchap_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def timing_enable(on):
    load().chap_timing_enable(1 if on else 0)


def timing_report():
    """{kernel family: dict(launches, ms, flops, bytes)} for the launches since timing_enable(True)."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(load().chap_timing_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, fl, by = line.split()
        out[name] = dict(launches=int(n), ms=float(ms), flops=float(fl), bytes=float(by))
    return out


def launch_count():
    return int(load().chap_launch_count())


def reset_launch_count():
    load().chap_reset_launch_count()
