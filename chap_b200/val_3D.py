"""code/val_3D.py of the reference (`test_single_case` :14-79, `cal_metric` :82-88, `test_all_case` :91-107) on the
sm_100a kernels.  test_single_case is the same function as in test_3D_util (the reference keeps two identical copies)."""
import numpy as np

from .test_3D_util import cal_metric, test_single_case  # noqa: F401


def test_all_case(net, cases, num_classes=4, patch_size=(48, 160, 160), stride_xy=32, stride_z=24, rank=0, world_size=1):
    """cases: iterable of (image, label) numpy volumes (the reference reads them from `.h5`, val_3D.py:99-102).
    Returns the mean [dice, hd95] per foreground class over this rank's cases (round-robin sharding, no collective)."""
    total, count = np.zeros((num_classes - 1, 2)), 0
    for i, (image, label) in enumerate(cases):
        if i % world_size != rank:
            continue
        prediction = test_single_case(net, image, stride_xy, stride_z, patch_size, num_classes=num_classes)
        for c in range(1, num_classes):
            total[c - 1, :] += cal_metric(label == c, prediction == c)
        count += 1
    return total / max(count, 1)
