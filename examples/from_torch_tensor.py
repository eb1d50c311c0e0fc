"""Counterpart of the reference's examples/python/basic/from_torch_tensor.py and to_torch_tensor.py: torch CUDA tensors in
and out through DLPack, without a copy (utility/dl_converter.h:32-40)."""
import numpy as np
import torch
from torch.utils.dlpack import from_dlpack, to_dlpack

from _clouds import pair
import cupoch_b200 as cph

if __name__ == "__main__":
    src, _, _, _, _ = pair(100_000)
    a = torch.from_numpy(src).cuda()
    pc = cph.geometry.PointCloud()
    pc.from_points_dlpack(to_dlpack(a))               # borrows the tensor's memory
    print("points in:", len(pc), "min bound", pc.get_min_bound(), "max bound", pc.get_max_bound())
    down = pc.voxel_down_sample(0.02)
    b = from_dlpack(down.to_points_dlpack())          # a torch view of the device vector
    print("down-sampled:", tuple(b.shape), b.device, "mean", b.mean(0).cpu().numpy())
    assert np.allclose(b.cpu().numpy(), down.points.cpu())
