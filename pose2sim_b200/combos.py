"""Host-side mirror of the kernel's candidate enumeration: k-subsets of the camera indices in
itertools.combinations (lexicographic) order, as bit masks (Pose2Sim/triangulation.py:411)."""
from math import comb

import numpy as np


def subset_masks(n, k):
    """uint32 masks of the k-subsets of {0..n-1}, lexicographic by sorted index tuple."""
    out = np.empty(comb(n, k), np.uint32)
    idx = list(range(k))
    i = 0
    while True:
        m = 0
        for c in idx:
            m |= 1 << c
        out[i] = m
        i += 1
        j = k - 1
        while j >= 0 and idx[j] == n - k + j:
            j -= 1
        if j < 0:
            break
        idx[j] += 1
        for t in range(j + 1, k):
            idx[t] = idx[t - 1] + 1
    return out


def unrank_subset(n, k, rank):
    """rank-th k-subset in the same order, without the table (what the kernel does past the table)."""
    mask, x = 0, 0
    for i in range(k):
        while True:
            cnt = comb(n - 1 - x, k - 1 - i)
            if rank < cnt:
                break
            rank -= cnt
            x += 1
        mask |= 1 << x
        x += 1
    return mask


def mask_to_ids(mask, n_cams):
    return [c for c in range(n_cams) if (int(mask) >> c) & 1]
