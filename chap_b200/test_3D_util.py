"""3D sliding-window inference on the sm_100a kernels.

Same entry points as the reference's code/test_3D_util.py (`test_single_case` :14-79,
`calculate_metric_percase` :147-152, `cal_dice` :132-144) and code/val_3D.py (`test_single_case`
:14-79, `cal_metric` :82-88).  The reference runs one window per forward, copies every window's
softmax to the host (8 MB at the LA configuration) and accumulates with numpy; here
  * windows are gathered on the device (`chap_sw_extract`) and run through the net in batches,
  * the logits of all windows stay in HBM (channels-last), and
  * ONE tile-owned, atomic-free kernel (`chap_sw_aggregate`) applies the softmax on load, adds the
    windows covering each voxel in the reference's x->y->z order (bit-identical fp32 sums), divides
    by the count and takes the first-max argmax.
Only the int64 label map (and optionally score / count maps) returns to the host.
"""
import math

import numpy as np
import torch

from . import ops


def _pad_amounts(shape, patch_size):
    pads = []
    for s, p in zip(shape, patch_size):
        tot = max(p - s, 0)
        pads.append((tot // 2, tot - tot // 2))
    return pads


def _net_logits(net, patches, output_index=0):
    y = net(patches)
    if isinstance(y, (tuple, list)):            # DualDecoder3d returns (o1, o2); test_LA.py:50 uses num_outputs=1
        return y[output_index]
    return y


def sliding_window_device(net, vol, stride_xy, stride_z, patch_size, num_classes, batch_windows=4, return_maps=False,
                          window_logits_hook=None):
    """Device-resident core of test_single_case: `vol` is a float32 CUDA tensor [W, H, D] already padded to at least the
    patch size; returns the int64 label volume on the device (and score [C, W, H, D] / count [W, H, D] with return_maps).
    Window loop, clamped last window and accumulation order follow code/test_3D_util.py:42-72."""
    ww, hh, dd = vol.shape
    sx = math.ceil((ww - patch_size[0]) / stride_xy) + 1
    sy = math.ceil((hh - patch_size[1]) / stride_xy) + 1
    sz = math.ceil((dd - patch_size[2]) / stride_z) + 1
    desc = ops.sw_desc((ww, hh, dd), patch_size, (sx, sy, sz), (stride_xy, stride_xy, stride_z), num_classes)
    n_win = sx * sy * sz
    pw, ph, pd = patch_size
    win = torch.empty((n_win, pw, ph, pd, num_classes), dtype=torch.float32, device=vol.device)   # channels-last logits
    was_training = net.training
    net.eval()
    with torch.no_grad():
        for first in range(0, n_win, batch_windows):
            count = min(batch_windows, n_win - first)
            patches = ops.sw_extract(desc, vol, first, count)
            logits = ops.cl(_net_logits(net, patches))
            if window_logits_hook is not None:
                window_logits_hook(first, logits)
            win[first:first + count].copy_(logits.permute(0, 2, 3, 4, 1))      # same memory order: plain copy
        out = ops.sw_aggregate(desc, win, is_prob=False, want_maps=return_maps)
    if was_training:
        net.train()
    return out


def test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=1, batch_windows=4,
                     device=None, return_maps=False, window_logits_hook=None):
    """image: numpy [W, H, D] float; returns label_map int64 numpy [W, H, D] (and, with
    return_maps, score_map [C, W, H, D] and cnt [W, H, D] float32) -- code/test_3D_util.py:14-79."""
    if device is None:
        device = next(net.parameters()).device
    w, h, d = image.shape
    pads = _pad_amounts(image.shape, patch_size)
    add_pad = any(a + b > 0 for a, b in pads)
    if add_pad:
        image = np.pad(image, pads, mode='constant', constant_values=0)
    vol = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(device)
    out = sliding_window_device(net, vol, stride_xy, stride_z, patch_size, num_classes, batch_windows, return_maps,
                                window_logits_hook)
    def labels_to_host(label):
        # int64 like np.argmax in the reference (:72); the class index crosses the bus as one byte per voxel and is widened on the host
        if num_classes <= 255:
            return label.to(torch.uint8).cpu().numpy().astype(np.int64)
        return label.cpu().numpy()
    if return_maps:
        label, score, cnt = out
        label_map, score_map, cnt_map = labels_to_host(label), score.cpu().numpy(), cnt.cpu().numpy()
    else:
        label_map = labels_to_host(out)
    if add_pad:
        (wl, _), (hl, _), (dl, _) = pads
        label_map = label_map[wl:wl + w, hl:hl + h, dl:dl + d]
        if return_maps:
            score_map = score_map[:, wl:wl + w, hl:hl + h, dl:dl + d]
            cnt_map = cnt_map[wl:wl + w, hl:hl + h, dl:dl + d]
    if return_maps:
        return label_map, score_map, cnt_map
    return label_map


def dice_coefficient(pred, gt):
    """medpy.metric.binary.dc: 2|A&B| / (|A| + |B|); 0.0 when both are empty."""
    pred, gt = np.asarray(pred).astype(bool), np.asarray(gt).astype(bool)
    size = np.count_nonzero(pred) + np.count_nonzero(gt)
    return 2.0 * np.count_nonzero(pred & gt) / float(size) if size > 0 else 0.0


def _surface_distances(a, b, spacing=None):
    from scipy import ndimage
    a, b = np.asarray(a).astype(bool), np.asarray(b).astype(bool)
    if not a.any() or not b.any():
        raise RuntimeError("surface distance is undefined for an empty mask")
    fp = ndimage.generate_binary_structure(a.ndim, 1)
    a_border = a ^ ndimage.binary_erosion(a, structure=fp, iterations=1)
    b_border = b ^ ndimage.binary_erosion(b, structure=fp, iterations=1)
    dt = ndimage.distance_transform_edt(~b_border, sampling=spacing)
    return dt[a_border]


def hd95(pred, gt, spacing=None):
    """medpy.metric.binary.hd95: 95th percentile of the symmetric surface distances."""
    return float(np.percentile(np.hstack((_surface_distances(pred, gt, spacing), _surface_distances(gt, pred, spacing))), 95))


def asd(pred, gt, spacing=None):
    """medpy.metric.binary.asd: average surface distance from pred to gt."""
    return float(_surface_distances(pred, gt, spacing).mean())


def jc(pred, gt):
    """medpy.metric.binary.jc: Jaccard coefficient."""
    pred, gt = np.asarray(pred).astype(bool), np.asarray(gt).astype(bool)
    union = np.count_nonzero(pred | gt)
    return float(np.count_nonzero(pred & gt)) / float(union) if union > 0 else 0.0


def cal_metric(gt, pred):
    """code/test_3D_util.py:82-88 / code/val_3D.py:82-88: [dice, hd95] or zeros."""
    gt, pred = np.asarray(gt), np.asarray(pred)
    if pred.sum() > 0 and gt.sum() > 0:
        return np.array([dice_coefficient(pred, gt), hd95(pred, gt)])
    return np.zeros(2)


def cal_dice(prediction, label, num=2):
    """code/test_3D_util.py:132-144 (np.float replaced by float)."""
    total_dice = np.zeros(num - 1)
    for i in range(1, num):
        p, l = (prediction == i).astype(float), (label == i).astype(float)
        total_dice[i - 1] += 2 * np.sum(p * l) / (np.sum(p) + np.sum(l))
    return total_dice


def ravd(pred, gt):
    """medpy.metric.binary.ravd: (|pred| - |gt|) / |gt| (raises on an empty reference like medpy does)."""
    pred, gt = np.asarray(pred).astype(bool), np.asarray(gt).astype(bool)
    v2 = np.count_nonzero(gt)
    if v2 == 0:
        raise RuntimeError("ravd is undefined for an empty reference mask")
    return (np.count_nonzero(pred) - v2) / float(v2)


def calculate_metric_percase(pred, gt):
    """code/test_3D_util.py:147-152: np.array([dice, |ravd|, hd95, asd]) (medpy restated on numpy / scipy).
    The reference lets medpy raise when a mask is empty; here an empty prediction or label gives the all-zero row
    (the convention of cal_metric, :82-88) instead of an exception in the middle of an evaluation run."""
    pred, gt = np.asarray(pred).astype(bool), np.asarray(gt).astype(bool)
    if not pred.any() or not gt.any():
        return np.zeros(4)
    return np.array([dice_coefficient(pred, gt), abs(ravd(pred, gt)), hd95(pred, gt), asd(pred, gt)])


def read_case(image_path):
    """(image, label) of one case.  The reference reads HDF5 datasets 'image' / 'label' (code/test_3D_util.py:101-103);
    h5py is used when it is installed, otherwise an `.npz` / `.npy` pair with the same stem is accepted (the on-disk
    layout is the same two arrays).  I/O only -- no device work."""
    stem = image_path[:-3] if image_path.endswith(".h5") else image_path
    import os
    if os.path.isfile(stem + ".h5"):
        try:
            import h5py
        except ImportError as e:                                   # pragma: no cover - depends on the image
            raise RuntimeError("%s.h5 exists but h5py is not installed; convert the case to %s.npz" % (stem, stem)) from e
        with h5py.File(stem + ".h5", "r") as f:
            return f["image"][:], f["label"][:]
    if os.path.isfile(stem + ".npz"):
        z = np.load(stem + ".npz")
        return z["image"], z["label"]
    raise FileNotFoundError("no case file %s.h5 / %s.npz" % (stem, stem))


def read_case_list(base_dir, test_list):
    """code/test_3D_util.py:92-95: one case id per line (text before the first comma) -> base_dir/data/<id>.h5"""
    with open(base_dir + '/{}'.format(test_list), 'r') as f:
        lines = f.readlines()
    return [base_dir + "/data/{}.h5".format(item.replace('\n', '').split(",")[0]) for item in lines if item.strip()]


def _sum_over_ranks(total, world_size):
    if world_size <= 1:
        return total
    import torch.distributed as dist
    if not dist.is_initialized():
        return total                                   # caller combines the partial means itself
    from .parallel import gather_rows
    return np.sum(np.stack(gather_rows(total, world_size)), axis=0)


def test_all_case(net, base_dir, method="unet_3D", test_list="full_test.list", num_classes=4, patch_size=(48, 160, 160),
                  stride_xy=32, stride_z=24, test_save_path=None, cases=None, batch_windows=4, rank=0, world_size=1):
    """code/test_3D_util.py:91-129, same positional signature and return value: total_metric / n_cases with
    total_metric [num_classes - 1, 4] holding [dice, |ravd|, hd95, asd] of class 1 in row 0 (the reference only fills
    row 0, :105-106).  Per-case lines go to `test_save_path/<method>.txt` like the reference; predictions are saved as
    `<id>_pred.npy` (the reference writes NIfTI through SimpleITK, which is I/O outside the hot path).
    Extensions (keyword only in practice): `cases` = in-memory list of (image, label) or (id, image, label) used instead
    of base_dir / test_list; `rank` / `world_size` shard the cases round-robin (replicas only) -- with an
    initialised process group the totals are summed over ranks, so every rank returns the global mean."""
    if cases is None:
        paths = read_case_list(base_dir, test_list)
        items = [(pth.split("/")[-1].replace(".h5", ""), pth) for pth in paths]
    else:
        items = [(c[0], c[1:]) if len(c) == 3 else ("case%d" % i, c) for i, c in enumerate(cases)]
    total_metric = np.zeros((num_classes - 1, 4))
    lines = []
    for i, (ids, src) in enumerate(items):
        if i % world_size != rank:
            continue
        image, label = read_case(src) if isinstance(src, str) else src
        prediction = test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=num_classes,
                                      batch_windows=batch_windows)
        metric = calculate_metric_percase(prediction == 1, label == 1)
        total_metric[0, :] += metric
        lines.append("{},{},{},{},{}\n".format(ids, metric[0], metric[1], metric[2], metric[3]))
        if test_save_path is not None:
            np.save(test_save_path + "/{}_pred.npy".format(ids), prediction.astype(np.uint8))
    total_metric = _sum_over_ranks(total_metric, world_size)
    n = max(len(items), 1)
    if test_save_path is not None:
        suffix = "" if world_size == 1 else ".rank%d" % rank
        with open(test_save_path + "/{}{}.txt".format(method, suffix), "a") as f:
            f.writelines(lines)
            f.writelines("Mean metrics,{},{},{},{}".format(*(total_metric[0] / n)))
    return total_metric / n
