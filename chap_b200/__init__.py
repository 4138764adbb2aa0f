"""chap_b200 -- B200-native (sm_100a) implementation of the CHAP training / inference hot path.

Host side: Python/PyTorch mirroring the reference's module API (networks, losses, evaluators);
device side: hand-written CUDA in chap_b200/csrc behind the C ABI of include/chap_b200.h.
"""
__version__ = "0.1.0"
