"""code/val_3D.py of the reference (`test_single_case` :14-79, `cal_metric` :82-88, `test_all_case` :91-107) on the
sm_100a kernels.  test_single_case is the same function as in test_3D_util (the reference keeps two identical copies)."""
import numpy as np

from .test_3D_util import _sum_over_ranks, cal_metric, read_case, read_case_list, test_single_case  # noqa: F401


def test_all_case(net, base_dir, test_list="full_test.list", num_classes=4, patch_size=(48, 160, 160), stride_xy=32,
                  stride_z=24, cases=None, batch_windows=4, rank=0, world_size=1):
    """code/val_3D.py:91-107, same positional signature: mean [dice, hd95] per foreground class, shape
    [num_classes - 1, 2], over the cases listed in base_dir/test_list (HDF5 'image' / 'label', :99-102).
    Extensions: `cases` = in-memory list of (image, label) instead of the files; `rank` / `world_size` shard the cases
    round-robin (replicas only, no data-path collective; with an initialised process group the per-class totals are
    summed over ranks so that every rank returns the global mean)."""
    items = read_case_list(base_dir, test_list) if cases is None else list(cases)
    total_metric = np.zeros((num_classes - 1, 2))
    for i, src in enumerate(items):
        if i % world_size != rank:
            continue
        image, label = read_case(src) if isinstance(src, str) else src
        prediction = test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=num_classes,
                                      batch_windows=batch_windows)
        for c in range(1, num_classes):
            total_metric[c - 1, :] += cal_metric(label == c, prediction == c)
    return _sum_over_ranks(total_metric, world_size) / max(len(items), 1)
