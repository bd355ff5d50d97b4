"""Loader for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _weights(npz):
    sd = {}
    for k in npz.files:
        if k.startswith("w/"):
            a = npz[k]
            sd[k[2:]] = torch.from_numpy(a.astype(np.float32)) if a.dtype == np.float16 else torch.from_numpy(a)
    return sd


def load_depthnet(mode: str):
    """Returns (state_dict, [stage dicts]) for mode 'adaptive' or 'variance'."""
    base = np.load(os.path.join(GOLDEN, "depthnet_adaptive.npz"))
    sd = _weights(base)
    npz = base
    if mode == "variance":
        npz = np.load(os.path.join(GOLDEN, "depthnet_variance.npz"))
        sd = {k: v for k, v in sd.items() if not k.startswith("DepthNet.")}
        sd.update(_weights(npz))
    stages = []
    for s in range(3):
        p = f"s{s}/"
        st = {k[len(p):]: torch.from_numpy(npz[k].astype(np.float32) if npz[k].dtype == np.float16 else npz[k])
              for k in npz.files if k.startswith(p)}
        st["features"] = list(st["features"].unbind(0))
        stages.append(st)
    return sd, stages


def load_homo():
    npz = np.load(os.path.join(GOLDEN, "homo_warping.npz"))
    return {k: torch.from_numpy(npz[k].astype(np.float32) if npz[k].dtype == np.float16 else npz[k]) for k in npz.files}
