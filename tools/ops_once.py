#!/usr/bin/env python3
"""Config 3 once, for ncu: VoxelDownSample(0.02) of 10 M points, index build over the down-sampled cloud,
SearchRadius(k=1, r=0.05) of the 10 M points (--self: against the 10 M cloud itself)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--self", action="store_true")
    ap.add_argument("--normals", type=int, default=0, help="also EstimateNormals(KNN k) of the down-sampled cloud")
    a = ap.parse_args()
    import cupoch_b200 as cph
    from cupoch_b200.testing import datagen
    p = datagen.uniform_cube(a.points, 21, hi=(4, 4, 1))
    pc = cph.geometry.PointCloud(p)
    down = pc.voxel_down_sample(0.02)
    tree = cph.geometry.KDTreeFlann(pc if a.self else down)
    cnt, idx, d2 = tree.search_radius(pc.points, 0.05, 1)
    if a.normals:
        down.estimate_normals(cph.geometry.KDTreeSearchParamKNN(a.normals))
    print("n_out", len(down), "found", cnt)


if __name__ == "__main__":
    main()
