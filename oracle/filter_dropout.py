"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product package.

Functional restatement of the APPLY half of perform_dropout (code/networks/FilterDropout.py:45-89) with the random
draws made explicit: `masks[idx]` is None (level untouched: `idx not in level`, :82-84) or the pair of per-(sample, channel)
factors the reference multiplies the unlabelled half with -- Binomial * 2 and its complement (:57-65), the Dropout2d(0.5)
noise (:67-68, 71-72) or the rescaled Bernoulli masks of scores_dropoutV2 / drop_based_on_prob (:74-78, 116-160).
Pinned against the imported reference in tests/test_oracle_vs_reference.py (masks recovered from the reference's own
outputs) and through tests/golden/filter_dropout.npz."""
import torch


def perform_dropout(x, masks):
    fp1, fp2 = [], []
    for feat, m in zip(x, masks):
        labeled_bs = feat.shape[0] // 2                                   # :52-53
        unlab = feat[labeled_bs:]
        if m is None:
            p1 = p2 = unlab
        else:
            shape = (unlab.shape[0], unlab.shape[1]) + (1,) * (feat.dim() - 2)
            p1, p2 = unlab * m[0].reshape(shape), unlab * m[1].reshape(shape)
        fp1.append(torch.cat((feat, p1)))                                 # :86-87
        fp2.append(torch.cat((feat, p2)))
    return fp1, fp2


def recover_masks(x, fp, level):
    """The factors a perform_dropout call applied, recovered from its output `fp` (one decoder's list): per (sample, channel)
    the ratio at the position of the largest |feature| value."""
    out = []
    for idx, (feat, f) in enumerate(zip(x, fp)):
        if idx not in level:
            out.append(None)
            continue
        n = feat.shape[0]
        unlab, pert = feat[n // 2:], f[n:]
        flat_u, flat_p = unlab.reshape(unlab.shape[0], unlab.shape[1], -1), pert.reshape(unlab.shape[0], unlab.shape[1], -1)
        k = flat_u.abs().argmax(dim=2, keepdim=True)
        out.append((flat_p.gather(2, k) / flat_u.gather(2, k)).squeeze(2))
    return out
