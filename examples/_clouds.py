"""Synthetic fragment pair for the examples (SURVEY 8d recipe), or the caller's own arrays (--npy src.npy tgt.npy)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cupoch_b200.testing import datagen  # noqa: E402


def pair(n=200_000, colors=False, extent=1.0):
    """-> (source points, target points, target colours or None, source colours or None, ground-truth 4x4)"""
    if "--npy" in sys.argv:
        i = sys.argv.index("--npy")
        return np.load(sys.argv[i + 1]).astype(np.float32), np.load(sys.argv[i + 2]).astype(np.float32), None, None, None
    tgt, _ = datagen.surface(n, 11, extent=extent)
    gt = datagen.gt_transform((-1.0, 1.5, 2.0), (0.01, -0.005, 0.008))
    if colors:
        tc = datagen.texture(tgt, 32, 0.01)
        src, sc = datagen.make_source(tgt, gt, 13, 14, 5e-4, attrs=[(tc, False)])
        return src, tgt, tc, sc, gt
    return datagen.make_source(tgt, gt, 13, 14, 5e-4), tgt, None, None, gt
