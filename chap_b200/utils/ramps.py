"""utils.ramps of the reference (absent; call site code/train_ours_2D.py:36).  Host scalars."""
import math


def sigmoid_rampup(current, rampup_length):
    """exp(-5 (1 - t)^2), t = clip(current, 0, L) / L  (arXiv:1610.02242, cited at train_ours_2D.py:35)."""
    if rampup_length == 0:
        return 1.0
    t = min(max(float(current), 0.0), float(rampup_length)) / float(rampup_length)
    return float(math.exp(-5.0 * (1.0 - t) ** 2))
