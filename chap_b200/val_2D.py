"""Per-volume 2D validation on the sm_100a kernels: same entry points as the reference's code/val_2D.py
(`test_single_volume` :54-97, `calculate_metric_percase` :43-51).

The reference runs the slices one by one (batch 1: zoom -> net -> softmax -> argmax -> .cpu() -> zoom back).  Here
all slices of the volume are zoomed on the host (scipy order-0, like the reference), stacked, and pushed through the
network in ONE forward in eval mode (BatchNorm uses running statistics, so samples are independent); the ensemble /
softmax / argmax run in one fused kernel; only the int64 label maps return to the host."""
import numpy as np
import torch
from scipy.ndimage import zoom

from . import ops
from .test_3D_util import dice_coefficient, hd95


def calculate_metric_percase(pred, gt):
    """code/val_2D.py:43-51 (medpy dc / hd95 restated on scipy)."""
    pred, gt = np.asarray(pred).copy(), np.asarray(gt).copy()
    pred[pred > 0] = 1
    gt[gt > 0] = 1
    if pred.sum() > 0:
        return dice_coefficient(pred, gt), (hd95(pred, gt) if gt.sum() > 0 else 0)
    return 0, 0


def predict_volume(image, net, patch_size=(256, 256), model_type='unet', device="cuda:0", max_batch=64):
    """int64 label volume [S, H, W] for a float volume [S, H, W] (numpy)."""
    s, x, y = image.shape
    stack = np.stack([zoom(image[i], (patch_size[0] / x, patch_size[1] / y), order=0) for i in range(s)])
    inp = torch.from_numpy(stack).unsqueeze(1).float().to(device)
    was_training = net.training
    net.eval()
    outs = []
    with torch.no_grad():
        for b0 in range(0, s, max_batch):
            o = net(inp[b0:b0 + max_batch])
            if model_type == 'model1':
                lab = ops.argmax(o[0])
            elif model_type == 'model2':
                lab = ops.argmax(o[1])
            elif model_type == 'logit_ensemble':                         # softmax((o1 + o2) / 2).argmax, val_2D.py:72-75
                lab = ops.argmax(o[0], o[1])
            elif model_type == 'prob_ensemble':                          # ((softmax o1 + softmax o2) / 2).argmax, :76-80
                lab = ops.argmax((ops.softmax(o[0]) + ops.softmax(o[1])).log())
            else:                                                        # single-output nets
                lab = ops.argmax(o[0] if isinstance(o, (tuple, list)) else o)
            outs.append(lab)
    if was_training:
        net.train()
    out = torch.cat(outs).cpu().numpy()
    return np.stack([zoom(out[i], (x / patch_size[0], y / patch_size[1]), order=0) for i in range(s)])


def test_single_volume(image, label, net, classes, patch_size=[256, 256], model_type='unet', device="cuda:0"):
    """image, label: tensors [1, S, H, W] (one validation volume); returns [(dice, hd95)] per foreground class."""
    image = image.squeeze(0).cpu().detach().numpy()
    label = label.squeeze(0).cpu().detach().numpy()
    prediction = predict_volume(image, net, tuple(patch_size), model_type, device)
    return [calculate_metric_percase(prediction == i, label == i) for i in range(1, classes)]
