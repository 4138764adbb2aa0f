"""utils.patch of the reference (absent; call site code/train_ours_2D.py:371)."""
from .. import ops


def create_maskV1(pseudo_outputs1, pseudo_outputs2, knowledge, scale_factor=4, topk=0.1):
    """Top-k patch mask over (mean knowledge + decoder disagreement); frozen spec in
    oracle/chap_losses.py create_mask_v1.  Returns float mask [N, *spatial]."""
    return ops.patch_topk_mask(knowledge, pseudo_outputs1, pseudo_outputs2, scale_factor, topk)
