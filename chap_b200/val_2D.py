"""Per-volume 2D validation on the sm_100a kernels: same entry points as the reference's code/val_2D.py
(`test_single_volume` :54-97, `calculate_metric_percase` :43-51).

The reference runs the slices one by one (batch 1: scipy zoom -> net -> softmax -> argmax -> .cpu() -> scipy zoom back).
Here the volume goes to the device once; all slices are resampled by a gather kernel that uses scipy.ndimage.zoom(order=0)'s
own index map (`ops.zoom_nearest`, bit-identical), pushed through the network in ONE forward per chunk in eval mode (BatchNorm
uses running statistics, so samples are independent), the ensemble / softmax / argmax run in one fused kernel, the label maps
are resampled back on the device and the per-class Dice sums come from one overlap-count kernel.  Only the int64 label volume
(for the HD95 surface distances, a host metric outside the hot path) and 3 * classes integers return to the host."""
import numpy as np
import torch

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


def predict_volume_device(image, net, patch_size=(256, 256), model_type='unet', max_batch=64):
    """int64 CUDA label volume [S, H, W] for a float32 CUDA volume [S, H, W]: zoom -> net -> argmax -> zoom back (:57-92)."""
    s, x, y = image.shape
    rx, ry = int(round(x * (patch_size[0] / x))), int(round(y * (patch_size[1] / y)))      # scipy's output shape: round(in * zoom)
    inp = ops.zoom_nearest(image.float(), rx, ry).unsqueeze(1)
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
    out = torch.cat(outs)
    bx, by = int(round(rx * (x / patch_size[0]))), int(round(ry * (y / patch_size[1])))
    return ops.zoom_nearest(out, bx, by)


def predict_volume(image, net, patch_size=(256, 256), model_type='unet', device="cuda:0", max_batch=64):
    """int64 label volume [S, H, W] for a float volume [S, H, W] (numpy in, numpy out)."""
    vol = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(device)
    return predict_volume_device(vol, net, tuple(patch_size), model_type, max_batch).cpu().numpy()


def test_single_volume(image, label, net, classes, patch_size=[256, 256], model_type='unet', device="cuda:0", with_hd95=True):
    """image, label: tensors [1, S, H, W] (one validation volume); returns [(dice, hd95)] per foreground class like
    code/val_2D.py:54-97.  Dice comes from the on-device overlap counts (2 |P & G| / (|P| + |G|), 0 when the prediction is
    empty, :46-51); HD95 needs the surfaces and is computed on the host from the returned label volume (with_hd95=False skips
    it and reports 0 -- the training loop only selects on Dice, code/train_ours_2D.py:417-421)."""
    vol = image.squeeze(0).detach().to(device=device, dtype=torch.float32)
    gt = label.squeeze(0).detach().to(device=device).to(torch.int64)
    prediction = predict_volume_device(vol, net, tuple(patch_size), model_type)
    counts = ops.label_overlap(prediction, gt, classes).cpu().numpy()
    pred_host = prediction.cpu().numpy() if with_hd95 else None
    gt_host = gt.cpu().numpy() if with_hd95 else None
    metric_list = []
    for i in range(1, classes):
        inter, n_pred, n_gt = (int(v) for v in counts[i])
        if n_pred == 0:
            metric_list.append((0, 0))
            continue
        dice = 2.0 * inter / float(n_pred + n_gt)
        hd = hd95(pred_host == i, gt_host == i) if (with_hd95 and n_gt > 0) else 0
        metric_list.append((dice, hd))
    return metric_list


def validate(valloader, net, classes, patch_size=[256, 256], model_type='logit_ensemble', device="cuda:0", rank=0, world_size=1,
             with_hd95=True):
    """The validation loop of code/train_ours_2D.py:407-415: mean [dice, hd95] per foreground class over the volumes of
    `valloader` (any iterable of {"image": [1, S, H, W], "label": [1, S, H, W]}).  Volumes are sharded round-robin over the ranks
    (replicas only); with an initialised process group the per-class sums are gathered so that every rank returns the global mean
    ("2D validation -- gather of per-class Dice")."""
    from .test_3D_util import _sum_over_ranks
    total, n = np.zeros((classes - 1, 2)), 0
    for i, batch in enumerate(valloader):
        n += 1
        if i % world_size != rank:
            continue
        total += np.array(test_single_volume(batch["image"], batch["label"], net, classes, patch_size, model_type, device, with_hd95),
                          dtype=np.float64).reshape(classes - 1, 2)
    return _sum_over_ranks(total, world_size) / max(n, 1)
