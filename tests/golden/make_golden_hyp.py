"""Golden fixture for the depth-hypothesis sampling neighbour of the path, produced by the
REFERENCE's own code: ``uncertainty_aware_samples`` (models/module.py:999-1038) between the two bilinear
up-samples and the trilinear resample exactly as models/cas_mvsnet.py:250-253, 269-274, 293-296 chain them.
Runs only in the build container.   Usage: python tests/golden/make_golden_hyp.py  (writes hypotheses.npz)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)


def main():
    from make_golden import load_reference
    cas, ref_module = load_reference()
    g = torch.Generator().manual_seed(21)
    B, H, W = 2, 64, 96
    blob = {}
    for name, (scale_prev, scale, D) in {"stage2": (4, 2, 32), "stage3": (2, 1, 8)}.items():
        hp, wp = H // scale_prev, W // scale_prev
        depth = 500 + 80 * torch.rand(B, hp, wp, generator=g)
        var = 0.5 + 6 * torch.rand(B, hp, wp, generator=g)
        var[0, 0, :4] = 900.0                      # exercises low_bound = -min(cur_depth, exp_var) = -cur_depth
        cur = F.interpolate(depth.unsqueeze(1), [H, W], mode="bilinear", align_corners=cas.Align_Corners_Range)
        ev = F.interpolate(var.unsqueeze(1), [H, W], mode="bilinear")
        full = ref_module.uncertainty_aware_samples(cur_depth=cur, exp_var=ev, ndepth=D, dtype=torch.float32,
                                                    device=torch.device("cpu"), shape=[B, H, W])
        out = F.interpolate(full.unsqueeze(1), [D, H // scale, W // scale], mode="trilinear",
                            align_corners=cas.Align_Corners_Range).squeeze(1)
        blob[name + "/depth"], blob[name + "/var"] = depth.numpy(), var.numpy()
        blob[name + "/full_sub"], blob[name + "/out"] = full[:, :, ::3, ::3].contiguous().numpy(), out.numpy()
        blob[name + "/meta"] = np.array([H, W, scale, D])
    dv = 425 + 2.65 * torch.arange(192, dtype=torch.float32).view(1, -1).repeat(B, 1)
    first = ref_module.uncertainty_aware_samples(cur_depth=dv, exp_var=None, ndepth=48, dtype=torch.float32,
                                                 device=torch.device("cpu"), shape=[B, H, W])
    blob["stage1/depth_values"], blob["stage1/out"] = dv.numpy(), first[:, :, ::16, ::16].numpy()
    np.savez_compressed(os.path.join(HERE, "hypotheses.npz"), **blob)
    print("done", os.path.getsize(os.path.join(HERE, "hypotheses.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
