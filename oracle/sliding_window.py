"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product package.

numpy restatement of the reference 3D sliding-window inference,
code/test_3D_util.py:14-79 (identical copy at code/val_3D.py:14-79), taking any callable
`net_fn(patch[1,1,pw,ph,pd] float32 ndarray) -> logits[1,C,pw,ph,pd] ndarray` so it runs
without CUDA (the reference hard-codes `.cuda()` at :59).

Order of floating point operations is the reference's: windows are visited x -> y -> z
and each window's softmax is ADDED to score_map in that order (:49-70); score_map is
divided by cnt BEFORE the argmax (:71-72); argmax ties resolve to the lowest class
(numpy semantics).
"""
import math

import numpy as np


def softmax_axis1(y):
    y = y - y.max(axis=1, keepdims=True)
    e = np.exp(y, dtype=np.float32)
    return e / e.sum(axis=1, keepdims=True)


def pad_amounts(shape, patch_size):
    """:17-36 -- symmetric zero padding up to the patch size (extra voxel on the right)."""
    pads = []
    for s, p in zip(shape, patch_size):
        tot = max(p - s, 0)
        pads.append((tot // 2, tot - tot // 2))
    return pads


def window_starts(dim, patch, stride):
    """:42-44,50-54 -- ceil((dim - patch)/stride) + 1 windows; the last one is clamped."""
    n = math.ceil((dim - patch) / stride) + 1
    return [min(stride * i, dim - patch) for i in range(n)]


def test_single_case(net_fn, image, stride_xy, stride_z, patch_size, num_classes=1,
                     softmax_fn=softmax_axis1, return_maps=False):
    w, h, d = image.shape
    pads = pad_amounts(image.shape, patch_size)
    add_pad = any(a + b > 0 for a, b in pads)
    if add_pad:
        image = np.pad(image, pads, mode="constant", constant_values=0)
    ww, hh, dd = image.shape
    xs_list = window_starts(ww, patch_size[0], stride_xy)
    ys_list = window_starts(hh, patch_size[1], stride_xy)
    zs_list = window_starts(dd, patch_size[2], stride_z)
    score_map = np.zeros((num_classes,) + image.shape, dtype=np.float32)
    cnt = np.zeros(image.shape, dtype=np.float32)
    for xs in xs_list:
        for ys in ys_list:
            for zs in zs_list:
                patch = image[xs:xs + patch_size[0], ys:ys + patch_size[1], zs:zs + patch_size[2]]
                patch = patch[None, None].astype(np.float32)
                y = softmax_fn(np.asarray(net_fn(patch), dtype=np.float32))[0]
                sl = (slice(xs, xs + patch_size[0]), slice(ys, ys + patch_size[1]), slice(zs, zs + patch_size[2]))
                score_map[(slice(None),) + sl] = score_map[(slice(None),) + sl] + y
                cnt[sl] = cnt[sl] + 1
    score_map = score_map / np.expand_dims(cnt, axis=0)
    label_map = np.argmax(score_map, axis=0)
    if add_pad:
        (wl, _), (hl, _), (dl, _) = pads
        label_map = label_map[wl:wl + w, hl:hl + h, dl:dl + d]
        score_map = score_map[:, wl:wl + w, hl:hl + h, dl:dl + d]
        cnt = cnt[wl:wl + w, hl:hl + h, dl:dl + d]
    if return_maps:
        return label_map, score_map, cnt
    return label_map
